#!/bin/bash
# A/B of the gradient exchange at N GPUs: bucket sizes / NCCL CTA caps.  usage: gpurun --gpus 4 -- 'bash tools/ddp_ab.sh 4 1024 r02bc "base b7 b7c16 b14"'
N=${1:-2}; GB=${2:-512}; TAG=${3:-ddp_ab}; WHICH=${4:-"base end_cta8 end_cta32 end_cta16 b50_cta16 base2"}
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
          "check", d.get("ddp_check", {}).get("grads_sum_ok"), d.get("ddp_check", {}).get("weights_identical"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
}
for w in $WHICH; do
  case $w in
    base|base2) run $w QV_DUMMY=1 ;;
    end_cta8)   run $w QV_DDP_BUCKET_MB=100000 QV_NCCL_MAX_CTAS=8 ;;
    end_cta16)  run $w QV_DDP_BUCKET_MB=100000 QV_NCCL_MAX_CTAS=16 ;;
    end_cta32)  run $w QV_DDP_BUCKET_MB=100000 QV_NCCL_MAX_CTAS=32 ;;
    b50_cta16)  run $w QV_DDP_BUCKET_MB=50 QV_NCCL_MAX_CTAS=16 ;;
    b7)         run $w QV_DDP_BUCKET_MB=7 ;;
    b7c16)      run $w QV_DDP_BUCKET_MB=7 QV_NCCL_MAX_CTAS=16 ;;
    b7c4)       run $w QV_DDP_BUCKET_MB=7 QV_NCCL_MAX_CTAS=4 ;;
    b14)        run $w QV_DDP_BUCKET_MB=14 ;;
    b14c16)     run $w QV_DDP_BUCKET_MB=14 QV_NCCL_MAX_CTAS=16 ;;
  esac
done
