#!/bin/bash
# Round-2 profiling captures (run on the GPU box, writes gpurun_out/): launch lists with DRAM bytes for the LM solve and both
# detectors, `ncu --set full` reports of the top kernels, SASS evidence of the FP64 tensor-core path, compute-sanitizer.
set -u
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/launches_lm_r02.csv python tools/lm_launches.py > /dev/null 2>&1
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/launches_detect_r02.csv python tools/detect_launches.py 214 aprilgrid > /dev/null 2>&1
ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/launches_chessboard_r02.csv python tools/detect_launches.py 102 chessboard > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:lm_accumulate -c 1 -o gpurun_out/ncu_lm_accumulate_r02 -f python tools/lm_launches.py > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"at_pre_kernel|at_merge_kernel|at_ccl_union" -c 3 -o gpurun_out/ncu_detect_r02 -f python tools/detect_launches.py 214 aprilgrid > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none -k regex:p3p_ransac -c 1 -o gpurun_out/ncu_p3p_r02 -f python -c "
import numpy as np, torch, sys
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
from ezcalib_b200 import lm, synth, ezcalibrator as E
from test_p3p_cpu import _views
ctx = lm.Context(0)
tgt = synth.aprilgrid_target(6, 6, 0.06, 0.3)
v, f, pp, W, H = _views(tgt, 2000, seed=25, W=3840, H=2160, f=2800.0)
c = np.stack(v)
E.compute_p3p_ransac_batch(ctx, c, tgt, f, pp, W, H)
torch.cuda.cudart().cudaProfilerStart(); E.compute_p3p_ransac_batch(ctx, c, tgt, f, pp, W, H); torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
" > /dev/null 2>&1
for r in ncu_lm_accumulate_r02 ncu_detect_r02 ncu_p3p_r02; do
  python tools/ncu_summary.py gpurun_out/$r.ncu-rep > gpurun_out/$r.txt 2>&1
  python tools/ncu_stalls.py gpurun_out/$r.ncu-rep >> gpurun_out/$r.txt 2>&1
done
{
  echo "## LM solve, radtan5 mono, 10,000,080 obs, 5 iterations (+3 no-op iterations of the last group)"; python tools/launch_summary.py gpurun_out/launches_lm_r02.csv
  echo; echo "## AprilGrid, 214 views 2448x2048 in flight"; python tools/launch_summary.py gpurun_out/launches_detect_r02.csv
  echo; echo "## Chessboard, 102 views 1280x1024 in flight"; python tools/launch_summary.py gpurun_out/launches_chessboard_r02.csv 20
} > gpurun_out/launch_shares_r02.txt 2>&1
python tools/ncu_traffic.py gpurun_out/launches_detect_r02.csv 214 > gpurun_out/traffic_detect_r02.txt 2>&1
python tools/ncu_traffic.py gpurun_out/launches_lm_r02.csv 1 > gpurun_out/traffic_lm_r02.txt 2>&1
cuobjdump -sass ezcalib_b200/libezcalib_b200.so 2>/dev/null | awk '/Function : .*lm_accumulate_kernel/{f=1} /Function : /{if(!/lm_accumulate_kernel/)f=0} f' > /tmp/sass_acc.txt
{ echo "# SASS of lm_accumulate_kernel instantiations (cuobjdump -sass libezcalib_b200.so): FP64 tensor-core opcodes"; echo "DMMA instructions: $(grep -c DMMA /tmp/sass_acc.txt)"; echo "DFMA instructions: $(grep -c DFMA /tmp/sass_acc.txt)"; grep -m 6 DMMA /tmp/sass_acc.txt; } > gpurun_out/sass_lm_accumulate_r02.txt
make sanitize > gpurun_out/sanitize_r02.log 2>&1; cp profiles/sanitize_*.txt gpurun_out/ 2>/dev/null
tail -3 gpurun_out/traffic_detect_r02.txt | head -2; head -12 gpurun_out/launch_shares_r02.txt; cat gpurun_out/sass_lm_accumulate_r02.txt | head -4; tail -2 gpurun_out/sanitize_memcheck.txt gpurun_out/sanitize_racecheck.txt gpurun_out/sanitize_synccheck.txt
