for cfg in "cg=1 share=0" "cg=2 share=0" "cg=2 share=1"; do for sh in F1 B2 F2p F2 B4; do python scripts/gemm_exp.py $sh $cfg; done; done
for cfg in "cg=2 share=1 bn=256" "cg=2 share=0 bn=256" ; do for sh in F2p F2 B4; do python scripts/gemm_exp.py $sh $cfg; done; done
for cfg in "cg=1" "cg=2" "cg=2 split=6" "cg=1 split=6"; do python scripts/gemm_exp.py W5 $cfg; python scripts/gemm_exp.py W3 $cfg; done
for r in 2 4; do KFP16_MMAREP=$r python scripts/gemm_exp.py F1 cg=2 share=0; done
