for m in 1 0; do for r in 0 1; do KFP16_MERGE=$m KFP16_NOROT=$r python scripts/gemm_exp.py W5 split=6; KFP16_MERGE=$m KFP16_NOROT=$r python scripts/gemm_exp.py W3 split=6; done; done
