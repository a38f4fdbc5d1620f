Q="--gpus 2 --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-bilateral --no-cupti"
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py $Q > gpurun_out/b2_$2.json 2> gpurun_out/b2_$2.err; rc=$?; echo "$2 rc=$rc"; tail -c 3000 gpurun_out/b2_$2.json | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['n_gpus'], d.get('loss_check'))" 2>/dev/null; return $rc; }
run 29511 default && exit 0
grep -E "Error|error" gpurun_out/b2_default.err | head -5
RSS_FUSE_ROW_STREAMS=0 run 29512 row0 && exit 0
RSS_STEM=0 run 29513 stem0 && exit 0
RSS_RES_LINK=0 run 29514 link0 && exit 0
RSS_FUSE_ROW_STREAMS=0 RSS_STEM=0 RSS_RES_LINK=0 run 29515 all0
