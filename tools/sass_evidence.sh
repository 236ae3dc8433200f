#!/bin/bash
# SASS evidence of the Blackwell-native path: per-kernel counts of the tcgen05 / TMA / TMEM mnemonics in libsom_b200.so.
#   bash tools/sass_evidence.sh > profiles/<tag>_sass.txt
SO=${1:-vit_som_b200/libsom_b200.so}
echo "# cuobjdump -sass $SO : mnemonics per kernel (tcgen05.mma -> UTC*MMA, TMA -> UTMALDG, tcgen05.ld/st -> LDTM/STTM, multimem.ld_reduce -> LDGMC, griddepcontrol.wait / launch_dependents -> ACQBULK / PREEXIT)"
cuobjdump -sass "$SO" | awk '
  /Function :/ { fn=$3; next }
  { if (match($0, /(UTC[A-Z]*MMA[A-Z0-9_.]*|UTMALDG[A-Z0-9_.]*|LDTM[A-Z0-9_.]*|STTM[A-Z0-9_.]*|UTCBAR[A-Z0-9_.]*|LDGMC[A-Z0-9_.]*|STGMC[A-Z0-9_.]*|ST[A-Z]*\.MC[A-Z0-9_.]*|ACQBULK|PREEXIT|UTCATOMSWS[A-Z0-9_.]*|USETMAXREG[A-Z0-9_.]*|UCGABAR_[A-Z]*|HMMA[A-Z0-9_.]*)/)) {
      m=substr($0, RSTART, RLENGTH); c[fn" "m]++ } }
  END { for (k in c) print c[k], k }' | sort -k2,2 -k1,1nr | awk '{printf "%6d  %-28s %s\n", $1, $3, $2}' | c++filt
