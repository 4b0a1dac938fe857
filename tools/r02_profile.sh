#!/bin/bash
# round-2 evidence run on one B200: ncu launch list of the bench command, ncu --set full of the attention / GEMM / LayerNorm
# kernels inside a ViT-L b64 step, compute-sanitizer on the tiny-model forward.  Summaries: tools/summarize_profiles.py r02
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export DINO_B200_GRAPH=0
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 24 -c 1 -f -o gpurun_out/r02_prof_attn \
    python tools/profile_step.py vitl14 64 2 > gpurun_out/r02_prof_attn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_f16 -s 98 -c 4 -f -o gpurun_out/r02_prof_gemm \
    python tools/profile_step.py vitl14 64 2 > gpurun_out/r02_prof_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:layernorm_kernel -s 49 -c 2 -f -o gpurun_out/r02_prof_ln \
    python tools/profile_step.py vitl14 64 2 > gpurun_out/r02_prof_ln.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02_sanitizer_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r02_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_sanitizer_synccheck.log 2>&1; echo "synccheck exit $?" >> gpurun_out/r02_sanitizer_synccheck.log
ls -la gpurun_out/r02_prof_*.ncu-rep gpurun_out/r02_launches_bench.csv; for f in gpurun_out/r02_sanitizer_*.log; do echo "== $f"; tail -n 4 "$f"; done
