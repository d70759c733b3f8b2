python scripts/gemm_exp.py G cg=2 m=256 n=160 bn=160 k=64 iters=30 prof=1
python scripts/gemm_exp.py G cg=1 m=256 n=160 bn=160 k=64 iters=30 prof=1
python scripts/gemm_exp.py G cg=2 m=9984 n=160 bn=160 k=64 iters=30 prof=1
for sh in F1 F2 B2 B4 W3 W5; do python scripts/gemm_exp.py $sh prof=1 iters=30; done
