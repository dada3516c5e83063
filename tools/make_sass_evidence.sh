#!/bin/bash
# profiles/r02_res_usage.txt and profiles/r02_sass_*: registers / stack / shared memory of every kernel of the headline
# shape, instruction mix and an excerpt of the dominant kernel's SASS (DMMA = mma.sync f64, UBLKCP = cp.async.bulk).
set -e
cd "$(dirname "$0")/.."
O=compressor-mpc_b200/obj/shape_coop_par.o
K='_ZN4cmpc15assemble_kernelINS_5ShapeILi0ELi3ELi2ELi2EEELi2ELi100EEEvNS_10StepParamsENS_11DeviceStateEPKd'
K200='_ZN4cmpc15assemble_kernelINS_5ShapeILi0ELi3ELi2ELi2EEELi4ELi200EEEvNS_10StepParamsENS_11DeviceStateEPKd'
{
  echo "# cuobjdump -res-usage compressor-mpc_b200/obj/shape_coop_par.o (Shape<0,3,2,2> = setup-coop-par), sm_100a"
  cuobjdump -res-usage $O | grep -A1 "Function" | sed 's/^ *//' | c++filt | sed 's/cmpc::StepParams, cmpc::DeviceState/.../'
} > profiles/r02_res_usage.txt
for pair in "p100:$K" "p200:$K200"; do
  tag=${pair%%:*}; fn=${pair#*:}
  cuobjdump -sass -fun "$fn" $O > /tmp/sass_$tag.txt
  {
    echo "# cuobjdump -sass -fun assemble_kernel<Shape<0,3,2,2>, ${tag}> compressor-mpc_b200/obj/shape_coop_par.o"
    echo "# instruction mix (mnemonic: count)"
    grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/sass_$tag.txt | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -28
    echo "# total instructions: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/sass_$tag.txt)   DMMA: $(grep -c 'DMMA' /tmp/sass_$tag.txt)   UBLKCP: $(grep -c 'UBLKCP' /tmp/sass_$tag.txt)   SYNCS (mbarrier): $(grep -c 'SYNCS' /tmp/sass_$tag.txt)"
    echo "# excerpt: the bulk copies of the load phase and the first DMMA block"
    grep -n -E "UBLKCP|SYNCS" /tmp/sass_$tag.txt | head -12
    first=$(grep -n "DMMA" /tmp/sass_$tag.txt | head -1 | cut -d: -f1)
    sed -n "$((first-6)),$((first+40))p" /tmp/sass_$tag.txt
  } > profiles/r02_sass_assemble_$tag.txt
done
wc -l profiles/r02_res_usage.txt profiles/r02_sass_assemble_p100.txt profiles/r02_sass_assemble_p200.txt
