for k in 64 3072; do python scripts/gemm_exp.py G cg=2 m=9984 n=160 bn=160 k=$k iters=20; done
for sh in F1 B2 F2 F2p B4 W3 W5; do python scripts/gemm_exp.py $sh; done
python scripts/gemm_exp.py F2 dbg=1 ctas_dbg=0,1
python scripts/gemm_exp.py B4 dbg=1 ctas_dbg=0
python scripts/gemm_exp.py F1 dbg=1 ctas_dbg=0
python scripts/gemm_exp.py W5 dbg=1 ctas_dbg=0 split=6
