# ncu evidence for the current kernels (1 GPU): launch list of the bench command, --set full of one layer's kernels
set -x
timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 160 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
timeout 60 python scripts/profile_tdnnf_layer.py 2 3 > gpurun_out/plain2.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"bn_relu_bwd|gemm_f16" -s 11 -c 11 -f -o gpurun_out/prof_r01_final python scripts/profile_tdnnf_layer.py 2 3 > gpurun_out/ncu2.log 2>&1
tail -n 2 gpurun_out/ncu.log gpurun_out/ncu2.log
