#!/bin/bash
# A/B of two builds of the product library on the SAME box: tools/bin/libu2gnn_b200_{old,new}.so are copied over the in-tree
# library in turn (old new old new); prints the FFN forward micro-benchmark and the default bench line's step time for each.
P=graph-transformer_b200/u2gnn_b200/libu2gnn_b200.so
for v in old new old new; do
  cp tools/bin/libu2gnn_b200_$v.so $P
  echo "== $v"
  python tools/bench_ffn_tc.py 4456448 128 1; python tools/bench_ffn_bwd.py 4456448 2>/dev/null | tail -2
  python bench.py --steps 30 --warmup 10 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', j['value'], j['ms_per_step'], j['roofline']['frac'], j['roofline']['avg_launch_ms'], j['clocks']['sm_mhz'])"
done
