for cg in 1 2; do python scripts/gemm_exp.py G cg=$cg m=256 n=160 bn=160 k=64 iters=30; python scripts/gemm_exp.py G cg=$cg m=9984 n=160 bn=160 k=64 iters=30; done
python scripts/gemm_exp.py G cg=2 m=9984 n=160 bn=160 k=64 iters=30 dbg=1 ctas_dbg=0 | grep "tile 7"
python scripts/elt_exp.py rot=1 | grep -E "pad|fold"
