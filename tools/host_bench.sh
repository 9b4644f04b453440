#!/bin/sh
# build + run the host micro-benchmark (optionally with gprof: PG=1)
set -e
cd "$(dirname "$0")/.."
FLAGS="-O2 -g -std=c++17 -I/usr/local/cuda/include -Iinclude -pthread"
[ -n "$PG" ] && FLAGS="$FLAGS -pg"
/usr/bin/g++ $FLAGS -o /tmp/host_bench tools/host_bench.cpp meilisearch_b200/csrc/host_index.cpp -Lcorpus -lindexgen -Lmeilisearch_b200 -lb200milli -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$(pwd)/corpus -Wl,-rpath,$(pwd)/meilisearch_b200
cd /tmp && ./host_bench "$@"
