#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
GEMM_D=384 GEMM_ROWS=43968 python tools/gemm_bench.py 2>&1 | grep "^gemm" | grep -v "+ln\|ln  "
GEMM_D=768 GEMM_ROWS=87680 python tools/gemm_bench.py 2>&1 | grep "^gemm" | grep -v "+ln\|ln  "
GEMM_D=384 GEMM_ROWS=43968 timeout 600 ncu --set full --clock-control none -k regex:gemm_f16_tcgen05 --launch-skip 3 -c 1 -o gpurun_out/r02_vits_qkv -f python tools/gemm_bench.py > gpurun_out/ncu_vits.log 2>&1
ncu -i gpurun_out/r02_vits_qkv.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active','sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_active.avg','smsp__cycles_active.avg','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']
for w in want:
    if w in h:
        i=h.index(w); print(w, rows[1][i], [r[i] for r in rows[2:]])
for i,c in enumerate(h):
    if 'tensor' in c and 'pct' in c: print(c, [r[i] for r in rows[2:]])
"
