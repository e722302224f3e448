#!/bin/bash
# Per-kernel counts of the Blackwell-only SASS instructions in librn_b200.so (tcgen05 = UTCIMMA/UTCBAR,
# TMA = UTMALDG, TMEM load = LDTM, FP64 tensor = DMMA) -- the committed proof that the hot path is sm_100a native.
SO=${1:-renormalizer_b200/librn_b200.so}
cuobjdump -sass "$SO" | awk '
  /Function :/ { fn=$3 }
  /UTCIMMA/ { c[fn,"UTCIMMA"]++ } /UTMALDG/ { c[fn,"UTMALDG"]++ } /LDTM/ { c[fn,"LDTM"]++ }
  /UTCBAR/ { c[fn,"UTCBAR"]++ } /DMMA/ { c[fn,"DMMA"]++ } /UTCATOMSWS|UTCCP/ { c[fn,"UTC_other"]++ }
  /SYNCS/ { c[fn,"SYNCS_mbarrier"]++ } /UCGABAR|CGABAR/ { c[fn,"CLUSTER_BAR"]++ }
  END { for (k in c) { split(k, a, SUBSEP); printf "%-8d %-18s %s\n", c[k], a[2], a[1] } }' | sort -k3,3 -k2,2
