# round-end measurement set (1 GPU).  Every command runs under its own timeout: a CUPTI trace of the PDL graphs has
# hung sporadically on the CNN workload (never without the profiler), so that trace is taken with KFP16_PDL=0 only.
set -x
timeout 200 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_err.txt
KFP16_WORKLOAD=cnn_tdnn timeout 200 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r01_cnn_final.json 2>> gpurun_out/bench_err.txt
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_reference.json 2>> gpurun_out/bench_err.txt
KFP16_PDL=0 timeout 120 python scripts/trace_step.py 16 gpurun_out/trace_final_nopdl.txt
timeout 120 python scripts/trace_step.py 16 gpurun_out/trace_final_pdl.txt
KFP16_PDL=0 timeout 120 python scripts/trace_step.py 16 gpurun_out/trace_final_cnn_nopdl.txt cnn_tdnn
