#!/bin/bash
# A/B of the gradient exchange at N GPUs (default 2): overlapped buckets vs one all-reduce at the end, NCCL CTA caps, SM margin.
# usage: gpurun --gpus 2 -- 'bash tools/ddp_ab.sh 2 512 r02aq'
N=${1:-2}; GB=${2:-512}; TAG=${3:-ddp_ab}
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" QV_BENCH_GLOBAL_BATCH=$GB python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - "$name" gpurun_out/${TAG}_${name}.json <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 3), "img/s", round(d["value"]), "ddp_overhead_ms", round(d.get("ddp_overhead_ms", -1), 3),
          "check", d.get("ddp_check", {}).get("grads_sum_exact"), d.get("ddp_check", {}).get("weights_identical"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
}
run base            QV_DUMMY=1
run end_cta8        QV_DDP_BUCKET_MB=100000 QV_NCCL_MAX_CTAS=8
run end_cta32       QV_DDP_BUCKET_MB=100000 QV_NCCL_MAX_CTAS=32
run end_cta16       QV_DDP_BUCKET_MB=100000 QV_NCCL_MAX_CTAS=16
run b50_cta16       QV_DDP_BUCKET_MB=50 QV_NCCL_MAX_CTAS=16
run base2           QV_DUMMY=1
