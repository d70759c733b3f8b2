# round-end measurement set (1 GPU): bench lines, in-graph trace, ncu launch list, ncu --set full of one layer's kernels
set -x
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_err.txt
KFP16_WORKLOAD=cnn_tdnn python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r01_cnn_final.json 2>> gpurun_out/bench_err.txt
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_reference.json 2>> gpurun_out/bench_err.txt
KFP16_PDL=0 python scripts/trace_step.py 16 gpurun_out/trace_final_nopdl.txt
python scripts/trace_step.py 16 gpurun_out/trace_final_pdl.txt
python scripts/trace_step.py 16 gpurun_out/trace_final_cnn_pdl.txt cnn_tdnn
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 200 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
python scripts/profile_tdnnf_layer.py 2 3 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"bn_relu_bwd|gemm_f16" -s 14 -c 14 -f -o gpurun_out/prof_r01_final python scripts/profile_tdnnf_layer.py 2 3 > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu.log gpurun_out/ncu2.log
