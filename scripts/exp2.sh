for r in 1 2 4; do KFP16_MMAREP=$r python scripts/gemm_exp.py F1 cg=1 share=0; done
for r in 1 2 4; do KFP16_MMAREP=$r python scripts/gemm_exp.py F1 cg=2 share=0; done
for r in 1 2 4; do KFP16_MMAREP=$r python scripts/gemm_exp.py F2p cg=1 bn=256; done
for r in 1 2 4; do KFP16_MMAREP=$r python scripts/gemm_exp.py B2 cg=1 share=0; done
