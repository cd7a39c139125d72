#!/bin/bash
# The gpurun calls behind the last numbers of round 2 (each line = the command of one call; outputs under gpurun_out/).
# differential fuzz campaigns (tests/fuzz_gpu.py): seeds 11 / 12 / 13, 150-200 s each
#   python tests/fuzz_gpu.py --seconds 150 --seed 11 ; python tests/fuzz_gpu.py --seconds 200 --seed 12 ; ... --seed 13
# where a load spends its time (before / after the table arena)
#   CRF_LOAD_TRACE=1 python profiles/prof_load.py --workload sr --reps 2
#   CRF_LOAD_TRACE=1 python profiles/prof_load.py --workload s22 --reps 3
# tile-shape variants on the small workloads
#   for v in "--wpt 8" "--wpt 16" "--flags 8" "--wpt 4"; do python profiles/prof_scan.py --workload sr --scale 1 --reps 3 $v; done
#   for v in "--wpt 8" "--wpt 16" "--flags 8"; do python profiles/prof_scan.py --workload s22 --scale 1 --reps 3 $v; done
# overlapped gather: A/B at N = 2 in one box session (gpurun --gpus 2), then N = 8 and N = 4 (gpurun --gpus 8)
#   for ov in 0 1; do CRF_XCHG_OVERLAP=$ov python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
#       --master-port 2951$ov bench.py --gpus 2; done
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 4
# final confirmation on the committed tree
python -m pytest tests -x -q -m gpu                     # 95 passed, 2 skipped
python -c "import __graft_entry__ as g; g.smoke()"      # smoke ok: 159164 bp, 380 repeats, bit-exact vs oracle
python bench.py                                         # 1134 Gbp/s, e2e 19.1 ms, roofline.frac 0.606
python bench.py --workload sr                           # 1173 Gbp/s, e2e 27.3 ms
