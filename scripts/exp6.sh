for k in 64 256 1024 3072 6144; do python scripts/gemm_exp.py G cg=2 m=9984 n=160 bn=160 k=$k iters=20; done
for k in 64 256 1024 3072 6144; do python scripts/gemm_exp.py G cg=1 m=9984 n=160 bn=160 k=$k iters=20; done
for k in 64 320 1024 3072; do python scripts/gemm_exp.py G cg=2 m=9984 n=1536 bn=256 k=$k iters=20; done
for k in 64 320 1024 3072; do python scripts/gemm_exp.py G cg=2 m=9984 n=1536 bn=128 k=$k iters=20; done
for k in 64 3072; do python scripts/gemm_exp.py G cg=2 m=9984 n=160 bn=160 k=$k iters=20 bk=1; done
