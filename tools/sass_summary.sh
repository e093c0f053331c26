#!/bin/bash
# Tensor-core / TMA / TMEM / mbarrier SASS mnemonics per kernel of libsgb200.so (proof that the hot contractions are
# tcgen05 + TMA code: UTCHMMA(.2CTA) = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA load / store).
# usage: tools/sass_summary.sh > profiles/r2_sass_conv_tc.txt
cd "$(dirname "$0")/.."
echo "# cuobjdump -sass speakerguard_b200/libsgb200.so: count, kernel, mnemonic (tools/sass_summary.sh)"
cuobjdump -sass speakerguard_b200/libsgb200.so | awk '/Function :/{fn=$3} /UTCHMMA|UTCQMMA|UTCMMA|UTMALDG|UTMASTG|UTMAPF|LDTM|STTM|UTCBAR|UTCCP|UTCATOM|SYNCS|UTMACCTL|UTMACMDFLUSH|UBLKCP|HMMA|MATCH/{ m=$0; sub(/^[ \t]*\/\*[0-9a-f]+\*\/[ \t]*/,"",m); split(m,a," "); op=a[1]; if (op ~ /^@/) op=a[2]; gsub(/;/,"",op); c[fn" "op]++} END{for(k in c) print c[k], k}' | sort -k2,2 -k1,1nr
