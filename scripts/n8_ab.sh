#!/bin/bash
# 8-GPU A/B of the NCCL CTA cap (SARSSL_NCCL_MAX_CTAS: 0 = NCCL default, 4 = library default) and the 1-GPU number on the same box.
cd "$(dirname "$0")/.."
for c in 0 4; do
  SARSSL_NCCL_MAX_CTAS=$c timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$c \
    bench.py --gpus 8 --steps 10 --warmup 3 --no-torch-eager --no-other-configs --no-input-pipeline > gpurun_out/n8_ctas$c.json 2> gpurun_out/n8_ctas$c.err
done
timeout 200 python bench.py --steps 10 --no-torch-eager --no-other-configs --no-input-pipeline --no-cpu-baseline > gpurun_out/n8_box_n1.json 2>/dev/null
echo done
