python scripts/gemm_exp.py W5 dbg=1 ctas_dbg=0,1 split=6
KFP16_MERGE=0 python scripts/gemm_exp.py W5 dbg=1 ctas_dbg=0 split=6
python scripts/gemm_exp.py W3 split=6
KFP16_MERGE=0 python scripts/gemm_exp.py W3 split=6
