for cg in 1 2; do for n in 64 128 160 256; do for r in 4 8; do KFP16_MMAREP=$r python scripts/gemm_exp.py G cg=$cg n=$n bn=$n k=2048 iters=10; done; done; done
for cg in 1 2; do for n in 160 256; do for r in 4 8; do KFP16_MMAREP=$r python scripts/gemm_exp.py G cg=$cg n=$n bn=$n k=2048 bk=1 iters=10; done; done; done
