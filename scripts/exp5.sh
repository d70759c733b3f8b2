timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x 2>&1 | tail -4
for cfg in "cg=1 share=0" "cg=2 share=0" "cg=2 share=1"; do for sh in F1 B2 F2p F2 B4; do python scripts/gemm_exp.py $sh $cfg; done; done
for cfg in "cg=2 share=1 bn=256" "cg=2 share=0 bn=256" "cg=1 share=0 bn=256"; do for sh in F2p F2 B4; do python scripts/gemm_exp.py $sh $cfg; done; done
for cfg in "cg=1" "cg=2" "cg=2 split=6" "cg=1 split=6"; do python scripts/gemm_exp.py W5 $cfg; python scripts/gemm_exp.py W3 $cfg; done
for n in 4096 8192; do for cg in 1 2; do python scripts/gemm_exp.py G cg=$cg m=$n n=$n k=$n iters=5; done; done
