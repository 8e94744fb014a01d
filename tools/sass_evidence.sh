#!/bin/bash
# SASS evidence for profiles/: which Blackwell-native instructions the shipped library holds, per kernel.
# usage: tools/sass_evidence.sh > profiles/r02_sass_evidence.txt
so=dbot_ros_b200/libdbot_b200.so
echo "# cuobjdump -sass $so (sm_100a), $(date -u +%Y-%m-%d), git $(git rev-parse --short HEAD)"
echo "# mnemonic counts over the whole library"
cuobjdump -sass $so | grep -oE "\b(UBLKCP|UBLKPF(\.L2)?|SYNCS(\.[A-Z_.]+)?|ATOMS(\.[A-Z0-9_.]+)?|REDS?(\.[A-Z0-9_.]+)?|CREDUX(\.[A-Z0-9_.]+)?|LDG\.E\.[A-Z0-9.]*256|STG\.E\.[A-Z0-9.]*256|LDG\.E\.128\.CONSTANT|SHFL\.(IDX|UP|DOWN|BFLY)|USETMAXREG[A-Z_.]*|BAR\.(SYNC|ARV)[A-Z_.]*|MUFU\.(RCP|EX2|LG2)|IMAD\.HI\.U32|VIMNMX3?|F2I\.(CEIL|FLOOR)[A-Z.]*|FENCE[A-Z_.]*|MEMBAR[A-Z_.]*)\b" | sort | uniq -c | sort -rn
echo
echo "# per particle kernel: TMA bulk copy (UBLKCP), mbarrier wait (SYNCS), L2 prefetch (UBLKPF), shared atomics, 256-bit global access"
for k in ILi0ELb0 ILi0ELb1 ILi1ELb0 ILi1ELb1 ILi2ELb0 ILi2ELb1; do
  cuobjdump -sass $so | awk -v k="dbot_particle_kernel$k" '/Function :/{p=index($0,k)>0} p' > /tmp/sass_$k.txt
  printf "dbot_particle_kernel<%s>: %d instructions" "$k" "$(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/sass_$k.txt)"
  for m in UBLKCP UBLKPF SYNCS "ATOMS.MAX" "ATOMS.OR" "ATOMS.ADD" CREDUX "LDG.E.NA.ENL2.256" "STG.E.NA.EFL2.256" SHFL "IMAD.HI.U32" USETMAXREG; do
    printf "  %s=%d" "$(echo $m | cut -d'\' -f1)" "$(grep -c "$m" /tmp/sass_$k.txt)"
  done
  echo
done
echo
echo "# the TMA staging of the mesh chunk's triangle list and the wait on its mbarrier (update kernel, single chunk)"
grep -n "UBLKCP\|SYNCS\|UBLKPF" /tmp/sass_ILi1ELb0.txt | cut -c1-140 | head -12
echo
echo "# the span unit's pixel loop (shared-memory max on the bits of the inverse depth) and the warp-wide trip count"
grep -n "CREDUX.MAX\|ATOMS.MAX" /tmp/sass_ILi0ELb0.txt | cut -c1-120 | head -12
