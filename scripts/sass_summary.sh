#!/bin/bash
# SASS evidence for the built library: tcgen05 / TMA / TMEM mnemonic counts, library dependencies, kernel inventory.
# usage: scripts/sass_summary.sh > profiles/rNN_sass_summary.txt
so=kaldi_fp16_b200/libkaldi_fp16.so
tmp=$(mktemp)
cuobjdump -sass $so > $tmp 2>/dev/null
echo "# $so  ($(stat -c %s $so) bytes), cuobjdump -sass | grep -c <mnemonic>"
for m in "UTCHMMA" "UTCHMMA.2CTA" "UTMALDG" "UTMASTG" "UTMAPF" "LDTM" "UTCBAR" "SYNCS" "FFMA2" "REDG.E.ADD.F32" "[^C]HMMA" "IMMA"; do
  printf "%-18s %s\n" "$m" "$(grep -c "$m" $tmp)"
done
echo "# (UTCHMMA = tcgen05.mma kind::f16, .2CTA = cta_group::2; UTMALDG/UTMASTG = cp.async.bulk.tensor load/store; LDTM = tcgen05.ld;"
echo "#  UTCBAR = tcgen05.commit; [^C]HMMA = mma.sync, none expected)"
echo "# ldd (no cuBLAS / cuDNN / NCCL on the product path):"
ldd $so | grep -v "linux-vdso\|ld-linux" | awk '{print "  " $1}'
echo "# kernels (cuobjdump 'Function :' demangled, template instances counted):"
grep "Function :" $tmp | sed 's/.*Function : //' | c++filt | sed 's/(.*//; s/<.*//' | sort | uniq -c | sort -rn
echo "# gemm_f16_sm100<BN, A_MN, B_MN, EpiKind, CG, MODE> instances:"
grep "Function :" $tmp | sed 's/.*Function : //' | c++filt | grep -o "gemm_f16_sm100<[^>]*>" | sort | uniq | tr '\n' ' ' | fold -w 150
echo
rm -f $tmp
