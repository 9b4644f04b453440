#!/bin/sh
# developer experiment: host-side scaling of the keyword step loop under allocator tunables and socket pinning (cfg 3)
N0=$(lscpu | sed -n 's/^NUMA node0 CPU(s): *//p')
N1=$(lscpu | sed -n 's/^NUMA node1 CPU(s): *//p')
echo "node0: $N0  node1: $N1  gpu numa: $(cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c | tr '\n' ' ')"
nvidia-smi topo -m 2>/dev/null | head -4
export CFGS=${CFGS:-4:1:32}
echo "== baseline"; python tools/keyword_e2e.py 2>&1 | tail -2
echo "== malloc tunables"
MALLOC_TRIM_THRESHOLD_=1073741824 MALLOC_MMAP_THRESHOLD_=1073741824 MALLOC_TOP_PAD_=67108864 python tools/keyword_e2e.py 2>&1 | tail -2
echo "== node0 pinned"; taskset -c "$N0" python tools/keyword_e2e.py 2>&1 | tail -2
echo "== node1 pinned"; taskset -c "$N1" python tools/keyword_e2e.py 2>&1 | tail -2
echo "== node0 pinned + tunables, 32 and 64 threads"
CFGS=4:1:32,4:1:64 MALLOC_TRIM_THRESHOLD_=1073741824 MALLOC_MMAP_THRESHOLD_=1073741824 MALLOC_TOP_PAD_=67108864 taskset -c "$N0" python tools/keyword_e2e.py 2>&1 | tail -4
