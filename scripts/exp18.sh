for a in 0 1; do for sh in F2 F2p B4; do KFP16_ASTAT=$a python scripts/gemm_exp.py $sh prof=1 iters=30 2>&1 | grep CUPTI; done; done
