"""GPU checks of the two list kernels at the end of the training step (csrc/loss_optim.cu), through the C ABI, bit-exact against
plain PyTorch: rss_shadow_cl_refresh (channels-last bf16 copies of the k x k conv weights, configs/base/loveda.py step semantics are
untouched by it) and rss_accum_bf16_list (fp32 accumulation of the library's bf16 weight gradients into the flat gradient buffer)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

# (Cout, Cin, k): vector paths (Cin % 8 == 0), scalar paths (Cin = 3 stem, Cin = 2 gate conv 7x7), rows that fill a chunk alone
SHAPES = [(64, 3, 3), (32, 32, 3), (64, 64, 3), (128, 32, 3), (256, 256, 3), (1, 2, 7), (480, 40, 3), (5, 24, 3)]


def test_shadow_cl_refresh_bit_exact():
    from representationlearning_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    ws, table, rows, off, total, nrow, max_row = [], [], [0], 0, 0, 0, 1
    for cout, cin, k in SHAPES:
        n = cout * cin * k * k
        ws.append((off, (cout, cin, k, k), total))
        table.append([off, total, cin, k * k])
        off += (n + 3) // 4 * 4                        # FlatSGD keeps parameters 16-byte aligned
        total += (n + 7) // 8 * 8
        nrow += cout
        rows.append(nrow)
        max_row = max(max_row, cin * k * k)
    flat = torch.randn(off, device="cuda", generator=g)
    out = torch.full((total,), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.shadow_cl_refresh(flat, out, torch.tensor(table, dtype=torch.int64).cuda(), torch.tensor(rows, dtype=torch.int64).cuda(),
                          len(table), max_row)
    torch.cuda.synchronize()
    for o, shp, d in ws:
        n = shp[0] * shp[1] * shp[2] * shp[3]
        want = flat[o:o + n].view(shp).permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).flatten()
        assert torch.equal(out[d:d + n], want), shp


def test_accum_bf16_list_bit_exact():
    from representationlearning_b200 import conv
    g = torch.Generator(device="cuda").manual_seed(9)
    items, want = [], []
    for cout, cin, k in SHAPES:
        sink = torch.randn(cout, cin, k, k, device="cuda", generator=g)
        dw = (torch.randn(cout, cin, k, k, device="cuda", generator=g) * 3).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        want.append(sink + dw.float())
        items.append((sink, dw, cin, k * k))
    for n in (5, 4096, 4097, 70000):                   # same-order entries (1x1 convs): partial, exact and multi-chunk
        sink = torch.randn(n, device="cuda", generator=g)
        dw = torch.randn(n, device="cuda", generator=g).to(torch.bfloat16)
        want.append(sink + dw.float())
        items.append((sink, dw, 1, 0))
    conv.PENDING["items"] = list(items)
    conv._flush_pending()
    torch.cuda.synchronize()
    for (sink, dw, cin, kk), w in zip(items, want):
        assert torch.equal(sink, w), (tuple(sink.shape), kk)


def test_bottleneck_residual_link_equivalence(monkeypatch):
    """_hrnet_rssformer.py:249-287 Bottleneck with identity residual: the gradient of the block input is d(conv1 path) + d(residual).
    With the hand-off (hrnet.RES_LINK) conv1's data gradient is one GEMM with beta = 1 into bn3's residual gradient; without it
    autograd adds the two tensors.  Same result up to one bf16 rounding."""
    from representationlearning_b200 import hrnet
    torch.manual_seed(11)
    blk = hrnet.Bottleneck(256, 64).cuda().train()
    x0 = torch.randn(2, 256, 24, 40, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    dy = torch.randn(2, 256, 24, 40, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    res = {}
    for on in (False, True):
        monkeypatch.setattr(hrnet, "RES_LINK", on)
        for p in blk.parameters():
            p.grad = None
        x = x0.clone().requires_grad_(True)
        y = blk(x)
        y.backward(dy)
        torch.cuda.synchronize()
        res[on] = (y.detach().float(), x.grad.float(), blk.conv1.weight.grad.clone(), blk.bn3.weight.grad.clone())
    # (the two runs are not bit-identical even in the forward pass: the BatchNorm sums are float atomics)
    def close(a, b, tol):
        return (a - b).norm().item() <= tol * b.norm().item()
    assert close(res[True][0], res[False][0], 1e-2)
    assert close(res[True][2], res[False][2], 2e-2) and close(res[True][3], res[False][3], 2e-2)
    a, b = res[True][1], res[False][1]
    assert close(a, b, 1e-2), (a - b).norm().item() / b.norm().item()      # bf16 rounding noise; a missing residual term would be O(1)
