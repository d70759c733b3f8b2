for sh in 0 1; do for r in 1 8; do KFP16_ASTAT=0 KFP16_MMAREP=$r python scripts/gemm_exp.py F2p cg=2 share=$sh iters=10; done; done
for sh in 0 1; do for r in 1 8; do KFP16_MMAREP=$r python scripts/gemm_exp.py F1 cg=2 share=$sh iters=10; done; done
