#!/usr/bin/env bash
# SASS mnemonic counts per object: evidence of the hardware paths each kernel file uses.
#   bash tools/sass_mnemonics.sh > profiles/rN_sass_mnemonics.txt      (after __graft_entry__.build(); no GPU needed)
echo "# SASS mnemonic counts per object (cuobjdump -sass vit-vs-raw-iq_b200/build/*.o, sm_100a): evidence of the hardware paths used"
echo "# UTCHMMA = tcgen05.mma, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = bulk (1-D TMA) copy, LDTM/STTM = tcgen05.ld/st,"
echo "# UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = mma.sync, LDSM = ldmatrix, MOVM = movmatrix, REDG = red.global"
for o in gemm_tc frontend_tc attn_tc5 attn_tiles attn_long attention attn_cls embed_smallk rowops gemm_simt; do
  f=vit-vs-raw-iq_b200/build/$o.o
  [ -f "$f" ] || continue
  echo "== $o.o"
  cuobjdump -sass "$f" | grep -oE "\b(UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|UBLKCP|LDTM|STTM|UTCBAR|UTCATOMSWS|UTMACCTL|UTMACMDFLUSH|SYNCS|HMMA|LDSM|MOVM|REDG|ATOMS|MUFU|FFMA|ELECT)\b" \
    | sort | uniq -c | sort -rn | awk '{printf "%7d %-10s", $1, $2} END {print ""}'
done
