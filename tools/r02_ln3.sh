#!/bin/bash
cd "$(dirname "$0")/.."
L=$PWD/dinov2.cpp_b200/lib
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "fused_layernorm" 2>&1 | tail -1
python tools/gemm_bench.py $L/libdinov2_b200.so 2>&1 | grep "gemm " | grep "proj\|fc2\|ln "
python tools/ln_prof.py $L/libdinov2_b200_lnprof.so
timeout 600 ncu --set full --clock-control none -k regex:gemm_f16_tcgen05 -c 6 -o gpurun_out/r02_ln_fused -f python tools/ln_prof.py $L/libdinov2_b200.so > gpurun_out/ncu_ln.log 2>&1
ncu -i gpurun_out/r02_ln_fused.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','lts__t_sectors_srcunit_tex_op_read.sum','lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum','lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum']
idx=[h.index(w) for w in want if w in h]
for r in rows[2:]:
    print([ (h[i].split('__')[-1], r[i]) for i in idx])
"
