#!/bin/bash
# Developer tool: which Blackwell-native instructions the built objects contain (cuobjdump -sass of fasta-python_b200/lib/*.o).
# Usage: bash tools/sass_evidence.sh > profiles/r02_sass_mnemonics.txt
PAT='\b(UTCIMMA|UTCHMMA|UTCQMMA|LDTM|STTM|UTMALDG[.A-Z0-9_]*|UBLKCP[.A-Z0-9_]*|UCGABAR[.A-Z_]*|SYNCS[.A-Z0-9_]*|DMMA[.0-9A-Z_]*|UTCBAR[.A-Z0-9_]*|UTCATOMSWS[.A-Z0-9_]*|ELECT|DFMA|MUFU\.RCP64H|MUFU\.RSQ64H|ST\.E\.[A-Z0-9.]*STRONG\.SYS|LD\.E\.[A-Z0-9.]*STRONG\.SYS|LDG\.E\.[A-Z0-9.]*STRONG\.(GPU|SYS)|STG\.E\.[A-Z0-9.]*STRONG\.(GPU|SYS))\b'
echo "# cuobjdump -sass, sm_100a objects of libfasta_b200.so: counts of the instructions that prove the Blackwell-native paths"
echo "#   UTCIMMA = tcgen05.mma (int8), LDTM = tcgen05.ld (TMEM), UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk,"
echo "#   SYNCS = mbarrier, UCGABAR = cluster barrier, DMMA = fp64 mma.sync, STRONG.GPU / STRONG.SYS = scoped exchange loads / stores"
for o in ozaki_gemm dense_stream dense_gsweep dense_sweep tv_stencil resident_loop batched_gemm vector_kernels legacy_rng; do
  echo; echo "== $o.o"
  cuobjdump -sass fasta-python_b200/lib/$o.o 2>/dev/null | grep -oE "$PAT" | sort | uniq -c | sort -rn | head -14
done
echo; echo "== sample lines"
cuobjdump -sass fasta-python_b200/lib/ozaki_gemm.o | grep -E "UTCIMMA|LDTM|UTMALDG" | head -6
cuobjdump -sass fasta-python_b200/lib/dense_gsweep.o | grep -E "UBLKCP|STG.E.128.STRONG.GPU|LDG.E.128.STRONG.GPU" | head -4
cuobjdump -sass fasta-python_b200/lib/tv_stencil.o | grep -E "UBLKCP" | head -2
cuobjdump -sass fasta-python_b200/lib/vector_kernels.o | grep -E "STRONG.SYS" | head -4
