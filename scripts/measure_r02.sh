# round-2 measurement set (1 GPU): bench lines, in-graph CUPTI trace, ncu launch list + --set full captures.
# Every command runs under its own timeout.  Outputs land in gpurun_out/; the summaries are copied to profiles/.
set -x
mkdir -p gpurun_out
T=${TAG:-final}
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_n1_$T.json 2> gpurun_out/r02_bench_n1_$T.err
timeout 250 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_$T.json 2>> gpurun_out/r02_bench_n1_$T.err
timeout 300 python bench.py --workload tdnnf_stack --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r02_bench_n1_tdnnf_stack_$T.json 2>> gpurun_out/r02_bench_n1_$T.err
KFP16_PDL=0 timeout 200 python scripts/trace_step.py 16 gpurun_out/r02_trace_cnn_tdnn_nopdl_$T.txt cnn_tdnn chain > gpurun_out/trace.log 2>&1
timeout 200 python scripts/trace_step.py 16 gpurun_out/r02_trace_cnn_tdnn_pdl_$T.txt cnn_tdnn chain > gpurun_out/trace1.log 2>&1
KFP16_PDL=0 timeout 200 python scripts/trace_step.py 16 gpurun_out/r02_trace_tdnnf_stack_nopdl_$T.txt > gpurun_out/trace2.log 2>&1
# ncu launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 400 --csv --log-file gpurun_out/r02_ncu_launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
# --set full of one eager CNN-TDNN step's kernels (GEMMs incl. the implicit convolutions, elementwise passes, chain, patches)
timeout 100 python scripts/profile_cnn_tdnn_step.py 1 > gpurun_out/plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none -k regex:"gemm_f16|bn_relu_bwd|colsum|chain_loss|im2col|col2im|sgd_kernel|softmax" -c 140 -f -o /tmp/r02_prof_cnn_tdnn python scripts/profile_cnn_tdnn_step.py 1 > gpurun_out/ncu2.log 2>&1
python scripts/ncu_summary.py /tmp/r02_prof_cnn_tdnn.ncu-rep > gpurun_out/r02_ncu_cnn_tdnn_step_$T.txt 2> gpurun_out/ncu_summary.err
python scripts/ncu_traffic.py /tmp/r02_prof_cnn_tdnn.ncu-rep cnn_tdnn > gpurun_out/r02_ncu_traffic_cnn_tdnn.json 2>> gpurun_out/ncu_summary.err
# source-level capture of the three dominant GEMM variants (small report, kept)
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_f16_sm100<128, false, true, 3|gemm_f16_sm100<160, false, true, 1|gemm_f16_sm100<128, false, false, 4" -c 3 -f -o gpurun_out/r02_prof_tdnnf_gemms python scripts/profile_cnn_tdnn_step.py 1 > gpurun_out/ncu3.log 2>&1
ls -la /tmp/r02_prof_cnn_tdnn.ncu-rep gpurun_out/*.ncu-rep
tail -n 2 gpurun_out/ncu.log gpurun_out/ncu2.log gpurun_out/ncu3.log
