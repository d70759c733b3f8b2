# round-2 measurement set (1 GPU): bench lines, in-graph CUPTI trace, ncu launch list + --set full captures.
# Every command runs under its own timeout.  Outputs land in gpurun_out/; the summaries are copied to profiles/.
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2>> gpurun_out/r02_bench_n1.err
timeout 300 python bench.py --workload tdnnf_stack --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r02_bench_n1_tdnnf_stack.json 2>> gpurun_out/r02_bench_n1.err
KFP16_PDL=0 timeout 200 python scripts/trace_step.py 16 gpurun_out/r02_trace_cnn_tdnn_nopdl.txt cnn_tdnn chain > gpurun_out/trace.log 2>&1
KFP16_PDL=0 timeout 200 python scripts/trace_step.py 16 gpurun_out/r02_trace_tdnnf_stack_nopdl.txt > gpurun_out/trace2.log 2>&1
# ncu launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 400 --csv --log-file gpurun_out/r02_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
# --set full of one eager CNN-TDNN step's kernels (GEMMs incl. the implicit convolutions, elementwise passes, chain)
timeout 100 python scripts/profile_cnn_tdnn_step.py 1 > gpurun_out/plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_f16|bn_relu_bwd|chain_loss|im2col|col2im|sgd_kernel" -c 110 -f -o gpurun_out/r02_prof_cnn_tdnn python scripts/profile_cnn_tdnn_step.py 1 > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu.log gpurun_out/ncu2.log
