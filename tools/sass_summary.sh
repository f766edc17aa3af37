#!/bin/bash
# Per-kernel SASS instruction-mix evidence for libjblas_b200.so (runs on the CPU box: cuobjdump needs no GPU).
#   bash tools/sass_summary.sh > profiles/sass_summary.txt
# Counts, per kernel family, the mnemonics that prove the Blackwell-native paths (B200_PROFILING.md "What proves ..."):
#   DMMA.8x8x4 = mma.sync.m8n8k4.f64 (FP64 tensor pipe)    UTMALDG = cp.async.bulk.tensor (TMA)     UBLKCP = cp.async.bulk
#   UTCHMMA    = tcgen05.mma.kind::tf32                    LDTM    = tcgen05.ld (TMEM)              UTCBAR/UTCATOMSWS = tcgen05 commit/alloc
#   FFMA2      = fma.rn.f32x2                              DFMA/FFMA = SIMT pipes                   SYNCS = mbarrier, LDGSTS = cp.async
set -euo pipefail
SO="${1:-jblas/jl_b200/libjblas_b200.so}"
echo "# SASS instruction mix of $SO ($(stat -c %s "$SO") bytes), $(nvcc --version | tail -1)"
echo "# produced by tools/sass_summary.sh; kernels grouped by function-name prefix (template instantiations summed)"
cuobjdump -sass "$SO" | awk '
  /Function : / { name=$3; sub(/^_ZN2jb[0-9]*/, "", name); sub(/^_Z[0-9]*/, "", name);
                  fam=name; sub(/I.*/, "", fam); sub(/Pd.*|Pf.*|PT_.*/, "", fam); if (fam == "") fam=name; nk[fam]++; cur=fam; next }
  cur != "" && /^ +\/\*[0-9a-f]+\*\// {
      for (i = 2; i <= NF; i++) if ($i !~ /^@/ && $i !~ /^\/\*/) { op=$i; break }
      sub(/;$/, "", op); split(op, p, "."); m=p[1];
      if (op ~ /^DMMA/) m="DMMA.8x8x4"; if (op ~ /^UTMALDG/) m="UTMALDG"; if (op ~ /^SYNCS/) m="SYNCS";
      if (m ~ /^(DMMA.8x8x4|UTMALDG|UBLKCP|UTCHMMA|LDTM|UTCBAR|UTCATOMSWS|FFMA2|DFMA|FFMA|SYNCS|LDGSTS|HMMA)$/) c[cur, m]++;
      tot[cur]++ }
  END {
      split("DMMA.8x8x4 UTMALDG UBLKCP UTCHMMA LDTM UTCBAR UTCATOMSWS FFMA2 DFMA FFMA SYNCS LDGSTS HMMA", cols, " ");
      printf "%-34s %6s %9s", "kernel family", "insts", "SASS"; for (j = 1; j <= 13; j++) printf " %10s", cols[j]; printf "\n";
      for (f in nk) { printf "%-34s %6d %9d", f, nk[f], tot[f]; for (j = 1; j <= 13; j++) printf " %10d", c[f, cols[j]]; printf "\n" }
  }' | (read -r hdr; echo "$hdr"; sort)
