set -x
T=final
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_$T.log 2>&1; tail -4 gpurun_out/r02_pytest_$T.log
timeout 300 python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_n1_$T.json 2> gpurun_out/r02_bench_n1_$T.err
timeout 250 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_$T.json 2>> gpurun_out/r02_bench_n1_$T.err
timeout 300 python bench.py --workload tdnnf_stack --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/r02_bench_n1_tdnnf_stack_$T.json 2>> gpurun_out/r02_bench_n1_$T.err
KFP16_PDL=0 timeout 200 python scripts/trace_step.py 16 gpurun_out/r02_trace_cnn_tdnn_nopdl_$T.txt cnn_tdnn chain > gpurun_out/trace.log 2>&1
timeout 200 python scripts/trace_step.py 16 gpurun_out/r02_trace_cnn_tdnn_pdl_$T.txt cnn_tdnn chain > gpurun_out/trace1.log 2>&1
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_$T.log 2>&1; tail -2 gpurun_out/r02_smoke_$T.log
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 400 --csv --log-file gpurun_out/r02_ncu_launches_$T.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
cut -c1-260 gpurun_out/r02_bench_n1_$T.json
