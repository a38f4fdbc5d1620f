/* oracle/bilateral_oracle.c — TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatement of the reference's permutohedral-lattice bilateral filter (the only native/FFI component of
 * Rongtao-Xu/RepresentationLearning), written from the algorithm, not from the text, of
 *   SCD-AAAI2023/wrapper/bilateralfilter/bilateralfilter.cpp:4-21    feature vector (x, y, r, g, b) / sigma
 *   SCD-AAAI2023/wrapper/bilateralfilter/bilateralfilter.cpp:24-41   one class plane at a time through the lattice
 *   SCD-AAAI2023/wrapper/bilateralfilter/bilateralfilter.cpp:43-55   the batch loop (OpenMP over images)
 *   SCD-AAAI2023/wrapper/bilateralfilter/permutohedral.cpp:116-300   Permutohedral::init, the SSE build (x86-64 always has __SSE__)
 *   SCD-AAAI2023/wrapper/bilateralfilter/permutohedral.cpp:490-553   Permutohedral::compute (splat / blur / slice), value_size 1
 * It follows the reference's ROUNDING SEQUENCE (every float multiply and add separately rounded, the same operand order,
 * pixels splatted in raster order) so that results are bit-identical to the compiled reference, and it reproduces one
 * quirk that changes results: the SSE build embeds pixels four at a time and, when H*W is not a multiple of 4, also embeds
 * the 1-3 padding lanes (feature vector 0) and CREATES their lattice points (permutohedral.cpp:175-178, 262-270); those
 * points carry no signal but relay the blur.
 *
 * Pinned by tests/test_cpu_bilateral.py against oracle/_ref/libbilateralfilter_ref.so (the reference's own two .cpp files
 * compiled by oracle/build_bilateral_ref.py) and against tests/golden/bilateral_*.npz produced by that library.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this file's library.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC oracle/bilateral_oracle.c -o oracle/libbilateral_oracle.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define D 5          /* feature dimensions: x, y, r, g, b */
#define V (D + 1)    /* vertices of a lattice simplex */

typedef struct { int16_t c[D]; } lkey;

/* ---- the set of lattice points: open addressing, ids in order of first appearance ---- */
typedef struct {
    lkey* keys;       /* id -> key */
    int32_t* slots;   /* hash slot -> id or -1 */
    size_t cap;       /* power of two */
    int count;
} pointset;

static uint64_t mix_key(const lkey* k) {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < D; ++i) {
        h ^= (uint16_t)k->c[i];
        h *= 0xBF58476D1CE4E5B9ull;
        h ^= h >> 29;
    }
    return h;
}

static int pointset_init(pointset* s, size_t max_points) {
    s->cap = 64;
    while (s->cap < 2 * max_points + 2) s->cap <<= 1;
    s->keys = (lkey*)malloc(sizeof(lkey) * (max_points + 1));
    s->slots = (int32_t*)malloc(sizeof(int32_t) * s->cap);
    s->count = 0;
    if (!s->keys || !s->slots) return -1;
    memset(s->slots, 0xFF, sizeof(int32_t) * s->cap);
    return 0;
}

static void pointset_free(pointset* s) { free(s->keys); free(s->slots); }

static int pointset_lookup(pointset* s, const lkey* k, int create) {
    size_t h = (size_t)mix_key(k) & (s->cap - 1);
    for (;;) {
        int id = s->slots[h];
        if (id < 0) {
            if (!create) return -1;
            s->keys[s->count] = *k;
            s->slots[h] = s->count;
            return s->count++;
        }
        if (memcmp(&s->keys[id], k, sizeof(lkey)) == 0) return id;
        h = (h + 1) & (s->cap - 1);
    }
}

/* ---- embedding of one feature vector: the enclosing simplex (V lattice keys) and barycentric weights ----
 * permutohedral.cpp:149-159 (constants), 182-262 (per-feature arithmetic, SSE lanes are independent). */
typedef struct { float scale[D]; float inv_v, v; } embed_consts;

static void embed_consts_init(embed_consts* ec) {
    float inv_std_dev = (float)(sqrt(2.0 / 3.0) * (double)V);
    for (int i = 0; i < D; ++i) ec->scale[i] = (float)(1.0 / sqrt((double)((i + 2) * (i + 1))) * (double)inv_std_dev);
    ec->inv_v = 1.0f / (float)V;
    ec->v = (float)V;
}

static void embed(const float f[D], const embed_consts* ec, lkey keys[V], float bary[V]) {
    float el[V], base[V], rank[V], b[V + 1];
    /* elevate onto the hyperplane sum = 0 */
    float run = 0.0f;
    for (int j = D; j > 0; --j) {
        float cf = f[j - 1] * ec->scale[j - 1];
        float jc = (float)j * cf;
        el[j] = run - jc;
        run = run + cf;
    }
    el[0] = run;
    /* nearest remainder-0 point: round half to even (cvtps2dq under the default MXCSR) */
    float coord_sum = 0.0f;
    for (int i = 0; i < V; ++i) {
        float q = rintf(ec->inv_v * el[i]);
        base[i] = q * ec->v;
        coord_sum = coord_sum + q;
    }
    /* rank of every coordinate's residual (descending order of residual = ascending rank) */
    for (int i = 0; i < V; ++i) rank[i] = 0.0f;
    for (int i = 0; i < D; ++i) {
        float di = el[i] - base[i];
        for (int j = i + 1; j < V; ++j) {
            float dj = el[j] - base[j];
            float lt = (di < dj) ? 1.0f : 0.0f;
            rank[i] = rank[i] + lt;
            rank[j] = rank[j] + (1.0f - lt);
        }
    }
    /* off-plane correction */
    for (int i = 0; i < V; ++i) {
        rank[i] = rank[i] + coord_sum;
        float up = (rank[i] < 0.0f) ? ec->v : 0.0f;
        float dn = (rank[i] >= ec->v) ? ec->v : 0.0f;
        float adj = up - dn;
        rank[i] = rank[i] + adj;
        base[i] = base[i] + adj;
    }
    /* barycentric weights */
    for (int i = 0; i < V + 1; ++i) b[i] = 0.0f;
    for (int i = 0; i < V; ++i) {
        float t = (el[i] - base[i]) * ec->inv_v;
        int p = D - (int)rank[i];
        b[p] = b[p] + t;
        b[p + 1] = b[p + 1] - t;
    }
    b[0] = b[0] + (1.0f + b[V]);
    /* the V vertices: vertex r adds the canonical-simplex row r, indexed by rank */
    for (int r = 0; r < V; ++r) {
        for (int i = 0; i < D; ++i) {
            int rk = (int)rank[i];
            int canon = (rk <= D - r) ? r : r - V;
            keys[r].c[i] = (int16_t)(base[i] + (float)canon);
        }
        bary[r] = b[r];
    }
}

/* ---- one image: lattice + K class planes (bilateralfilter.cpp:24-41) ---- */
static int filter_image(const float* image, const float* in, float* out, int K, int H, int W, float sigmargb, float sigmaxy,
                        int* lattice_points) {
    const int P = H * W, P4 = (P + 3) & ~3;
    embed_consts ec;
    embed_consts_init(&ec);
    pointset ps;
    if (pointset_init(&ps, (size_t)P4 * V)) return -1;
    int32_t* vertex = (int32_t*)malloc(sizeof(int32_t) * (size_t)P4 * V);
    float* weight = (float*)malloc(sizeof(float) * (size_t)P4 * V);
    if (!vertex || !weight) return -1;
    for (int p = 0; p < P4; ++p) {
        float f[D] = {0, 0, 0, 0, 0};          /* padding lanes of the last SSE block: zero feature, points still created */
        if (p < P) {
            int x = p % W, y = p / W;
            f[0] = (float)x / sigmaxy;
            f[1] = (float)y / sigmaxy;
            f[2] = image[0 * P + p] / sigmargb;
            f[3] = image[1 * P + p] / sigmargb;
            f[4] = image[2 * P + p] / sigmargb;
        }
        lkey keys[V];
        float bary[V];
        embed(f, &ec, keys, bary);
        for (int r = 0; r < V; ++r) {
            vertex[(size_t)p * V + r] = pointset_lookup(&ps, &keys[r], 1);
            weight[(size_t)p * V + r] = bary[r];
        }
    }
    const int M = ps.count;
    if (lattice_points) *lattice_points = M;
    /* blur neighbours along each of the V lattice directions (permutohedral.cpp:279-299): +-1 on every stored coordinate,
       -+D on the direction's own coordinate (the last direction's own coordinate is the implicit one) */
    int32_t* nb = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)V * (M + 1));
    float* va = (float*)calloc((size_t)M + 2, sizeof(float));
    float* vb = (float*)calloc((size_t)M + 2, sizeof(float));
    if (!nb || !va || !vb) return -1;
    for (int dir = 0; dir < V; ++dir)
        for (int i = 0; i < M; ++i) {
            lkey lo = ps.keys[i], hi = ps.keys[i];
            for (int k = 0; k < D; ++k) { lo.c[k] = (int16_t)(lo.c[k] - 1); hi.c[k] = (int16_t)(hi.c[k] + 1); }
            if (dir < D) { lo.c[dir] = (int16_t)(ps.keys[i].c[dir] + D); hi.c[dir] = (int16_t)(ps.keys[i].c[dir] - D); }
            nb[((size_t)dir * M + i) * 2 + 0] = pointset_lookup(&ps, &lo, 0);
            nb[((size_t)dir * M + i) * 2 + 1] = pointset_lookup(&ps, &hi, 0);
        }
    const float alpha = 1.0f / (1.0f + powf(2.0f, -(float)D));
    for (int k = 0; k < K; ++k) {
        const float* src = in + (size_t)k * P;
        float* dst = out + (size_t)k * P;
        memset(va, 0, sizeof(float) * ((size_t)M + 2));
        memset(vb, 0, sizeof(float) * ((size_t)M + 2));
        /* splat, raster order; slot 0 stands for "no such lattice point" */
        for (int p = 0; p < P; ++p)
            for (int r = 0; r < V; ++r) {
                size_t o = (size_t)vertex[(size_t)p * V + r] + 1;
                float t = weight[(size_t)p * V + r] * src[p];
                va[o] = va[o] + t;
            }
        /* blur: one [1/2, 1, 1/2] pass per direction, Jacobi style */
        float *cur = va, *nxt = vb;
        for (int dir = 0; dir < V; ++dir) {
            for (int i = 0; i < M; ++i) {
                float a = cur[nb[((size_t)dir * M + i) * 2 + 0] + 1];
                float b = cur[nb[((size_t)dir * M + i) * 2 + 1] + 1];
                float s = a + b;
                float h = 0.5f * s;
                nxt[i + 1] = cur[i + 1] + h;
            }
            float* t = cur; cur = nxt; nxt = t;
        }
        /* slice */
        for (int p = 0; p < P; ++p) {
            float acc = 0.0f;
            for (int r = 0; r < V; ++r) {
                float w = weight[(size_t)p * V + r] * alpha;
                float t = w * cur[(size_t)vertex[(size_t)p * V + r] + 1];
                acc = acc + t;
            }
            dst[p] = acc;
        }
    }
    free(nb); free(va); free(vb); free(vertex); free(weight);
    pointset_free(&ps);
    return 0;
}

/* Same argument meaning as the reference's bilateralfilter_batch (bilateralfilter.hpp:12): images (N,3,H,W), ins/outs (N,K,H,W),
 * flat fp32, caller-allocated; returns 0 or -1 (allocation failure).  lattice_points: optional [N] out (M per image). */
int bilateral_oracle_batch(const float* images, const float* ins, float* outs, int N, int K, int H, int W, float sigmargb,
                           float sigmaxy, int* lattice_points) {
    int status = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int n = 0; n < N; ++n) {
        size_t P = (size_t)H * W;
        int rc = filter_image(images + (size_t)n * 3 * P, ins + (size_t)n * K * P, outs + (size_t)n * K * P, K, H, W, sigmargb,
                              sigmaxy, lattice_points ? lattice_points + n : 0);
        if (rc) {
#pragma omp atomic write
            status = rc;
        }
    }
    return status;
}
