for cfg in "cg=1 share=0" "cg=1 share=0 flush=1" "cg=2 share=0" "cg=2 share=1" "cg=1 share=0 ctas=39" "cg=1 share=0 pf=4" "cg=1 share=0 pf=8" "cg=1 share=0 pf=8 flush=1" "cg=2 share=1 pf=4" "cg=1 share=0 split=2"; do python scripts/gemm_exp.py F1 $cfg; done
for cfg in "cg=1" "cg=2" "cg=1 bn=256" "cg=2 bn=256" "cg=2 bn=256 share=0"; do python scripts/gemm_exp.py F2p $cfg; python scripts/gemm_exp.py F2 $cfg; done
for cfg in "cg=1" "cg=2" "cg=1 split=6" "cg=1 split=3"; do python scripts/gemm_exp.py W5 $cfg; python scripts/gemm_exp.py W3 $cfg; done
