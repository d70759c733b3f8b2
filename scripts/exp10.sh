for n in 128 256; do for r in 1 8; do KFP16_MMAREP=$r python scripts/gemm_exp.py G cg=2 m=9984 n=1536 bn=$n k=2048 iters=10; done; done
for n in 160; do for r in 1 8; do KFP16_MMAREP=$r python scripts/gemm_exp.py G cg=2 m=9984 n=160 bn=$n k=3072 iters=10; done; done
