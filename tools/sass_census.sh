#!/bin/bash
# SASS opcode census of the tensor-core objects: proof that the shipped kernels are tcgen05 / TMA / TMEM native.
# usage: tools/sass_census.sh > profiles/r02_sass_census.txt
cd "$(dirname "$0")/.."
echo "# cuobjdump -sass of unet_convlstm_b200/build/*.o (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a), instruction counts"
echo "# UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG = cp.async.bulk.tensor (TMA load), LDTM = tcgen05.ld,"
echo "# UTCBAR = tcgen05.commit, UTMAPF/UTMACCTL = tensormap prefetch, SYNCS = mbarrier ops, UTCATOMSWS = tcgen05.alloc/dealloc"
for o in conv_tc conv_tc2 conv_halo wgrad_tc wgrad_tc2 wgrad_halo; do
  f=unet_convlstm_b200/build/$o.o
  [ -f "$f" ] || continue
  echo; echo "== $o.o"
  cuobjdump -sass "$f" | grep -oE "\b(UTCHMMA(\.2CTA)?|UTCQMMA|UTMALDG(\.[0-9]D)?(\.2CTA)?|LDTM(\.x[0-9]+)?|STTM|UTCBAR(\.2CTA)?(\.MULTICAST)?|UTCATOMSWS[.A-Z0-9_]*|UTMAPF|SYNCS[.A-Z0-9_]*|ELECT|STG\.E\.ENL2\.256|RED\.E[.A-Z0-9_]*|HMMA[.0-9A-Z]*|MUFU\.TANH)\b" | sort | uniq -c | sort -rn
done
