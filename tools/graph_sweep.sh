#!/bin/bash
# priority / block-width sweep of the graph-replayed evaluation at n = 8192 (tuning aid; results in profiles/)
export GRAPH_NS=8192 GRAPH_MODES=1 GRAPH_STEPS=10
for prio in "" "-5,-4,0,0,0" "-5,-4,0,-2,-1" "-5,-3,-1,-2,0" "-5,-5,0,0,0" "-5,-4,-1,-3,-2"; do
  echo "PRIO=[$prio]"; GPK_GRAPH_PRIO="$prio" GPK_GRAPH_DEBUG=1 timeout 100 python tools/graph_check.py 2>&1 | grep -v "^$"
done
for nb in 256 384 640; do
  echo "NB=$nb"; GPK_PIPE_NB=$nb timeout 100 python tools/graph_check.py 2>&1 | grep "ms/eval"
done
