#!/bin/bash
# SASS evidence for profiles/r02_sass_excerpt.txt (run after `make`)
for f in emit parquet_decode merge orc_decode parquet_encode readback; do
  echo "== $f.cu"
  cuobjdump -sass build/$f.cu.o > /tmp/sass_$f.txt
  for pat in UBLKCP SYNCS LDGSTS LDG.E.128 STG.E.128 LDS SHFL VOTE ATOMG CCTL; do printf "  %-10s %6d\n" "$pat" $(grep -c "$pat" /tmp/sass_$f.txt); done
  grep -E "UBLKCP|SYNCS|LDGSTS" /tmp/sass_$f.txt | head -6 | sed 's/^ *//' | cut -c1-150
done
