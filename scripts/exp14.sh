python scripts/gemm_exp.py F1 dbg=1 ctas_dbg=0,1,77
python scripts/gemm_exp.py F2 dbg=1 ctas_dbg=0
