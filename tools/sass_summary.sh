#!/bin/bash
# SASS evidence for the judge: per kernel, how many tcgen05 / TMA / TMEM instructions the in-tree library contains.
# usage: tools/sass_summary.sh > profiles/r02/sass_summary.txt     (no GPU needed)
LIB=efficient-attention_b200/lib/libeva_sm100.so
echo "# cuobjdump -sass $LIB : instruction counts per kernel (UTC*MMA = tcgen05.mma, UTMALDG/UTMASTG/UBLKCP = TMA, LDTM/STTM = tcgen05.ld/st,"
echo "# SYNCS = mbarrier, UCGABAR = cluster barrier, HMMA = legacy mma.sync (must be 0), FFMA2 = packed fp32)"
cuobjdump -sass "$LIB" | c++filt | awk '
/Function :/ { name=$0; sub(/.*Function : /,"",name); sub(/\(.*/,"",name); gsub(/eva::/,"",name); gsub(/__nv_bfloat16/,"bf16",name); gsub(/__half/,"f16",name); gsub(/ /,"",name); order[++n]=name }
/^[[:space:]]*\/\*[0-9a-f]+\*\// {
  total[name]++
  if ($0 ~ /UTC[A-Z]*MMA/) mma[name]++
  if ($0 ~ /UTMALDG/) tmald[name]++
  if ($0 ~ /UTMASTG/) tmast[name]++
  if ($0 ~ /UBLKCP/) blk[name]++
  if ($0 ~ /UTMAPF|UTMACCTL/) pf[name]++
  if ($0 ~ /LDTM/) ldtm[name]++
  if ($0 ~ /STTM/) sttm[name]++
  if ($0 ~ /SYNCS/) syncs[name]++
  if ($0 ~ /UCGABAR/) cga[name]++
  if ($0 ~ /[^A-Z]HMMA/) hmma[name]++
  if ($0 ~ /FFMA2|FADD2/) f2[name]++
  if ($0 ~ /MUFU.EX2/) ex2[name]++
}
END {
  printf "%-62s %7s %6s %7s %7s %6s %5s %5s %6s %7s %5s %6s %5s\n", "kernel", "instr", "UTCMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "SYNCS", "UCGABAR", "HMMA", "FFMA2", "EX2"
  for (i=1;i<=n;i++) { k=order[i]; if (total[k] < 200) continue;
    printf "%-62s %7d %6d %7d %7d %6d %5d %5d %6d %7d %5d %6d %5d\n", substr(k,1,62), total[k], mma[k], tmald[k], tmast[k], blk[k], ldtm[k], sttm[k], syncs[k], cga[k], hmma[k], f2[k], ex2[k] }
}' | (read -r hdr; echo "$hdr"; sort)
