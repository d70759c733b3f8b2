for sh in F2 B4; do python scripts/gemm_exp.py $sh prof=1 iters=30 2>&1 | grep CUPTI; done
KFP16_PDL=0 timeout 200 python scripts/trace_step.py 16 gpurun_out/trace_ring5.txt 2>&1 | tail -1; grep "gemm_f16_sm100<128" gpurun_out/trace_ring5.txt | head -2
