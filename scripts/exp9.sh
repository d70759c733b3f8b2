for sh in F2 F2p B4; do python scripts/gemm_exp.py $sh; KFP16_ASTAT=0 python scripts/gemm_exp.py $sh; done
python scripts/gemm_exp.py F2 dbg=1 ctas_dbg=0
python scripts/gemm_exp.py B4 dbg=1 ctas_dbg=0
