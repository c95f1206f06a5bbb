#!/bin/bash
# SASS opcode evidence of the built library (runs without a GPU): tcgen05 / TMA / TMEM mnemonics per B200_PROFILING.md.
SO=ipr_gan_b200/libipr_b200.so
T=$(mktemp); cuobjdump -sass $SO > $T
echo "SASS opcode evidence: libipr_b200.so built for sm_100a (cuobjdump -sass, lines containing the mnemonic)"
for m in UTCHMMA UTCBAR UTMALDG.4D UTMALDG.2D LDTM SYNCS; do printf "%-22s %s\n" $m $(grep -c "$m" $T); done
printf "%-22s %s\n" "legacy HMMA (mma.sync)" $(grep -c " HMMA" $T)
printf "%-22s %s\n" "HGMMA (wgmma)" $(grep -c "HGMMA" $T)
echo
echo "tcgen05.mma -> UTCHMMA, tcgen05.commit -> UTCBAR, cp.async.bulk.tensor -> UTMALDG, tcgen05.ld -> LDTM, mbarrier -> SYNCS"
echo
echo "kernel entry points: $(grep -c 'Function :' $T)"
grep 'Function :' $T | sed -E 's/.*Function : //; s/_ZN[0-9]+_GLOBAL__N__[0-9a-f]+_[0-9]+_//; s/_cu_[0-9a-f]+/_cu_/' | sed -E 's/(tapgemm_kernel|wgrad_kernel).*/\1/' | sed -E 's/^([a-z_0-9]+_cu_)[0-9]*([a-z_0-9]+kernel).*/\1\2/' | sort | uniq -c | sort -rn | head -40
rm -f $T
