#!/bin/bash
for sk in 0 1 2 4 5 7; do
  echo "== MCN_DEBUG_SKIP=$sk"; MCN_DEBUG_SKIP=$sk build/tc_harness 34; MCN_DEBUG_SKIP=$sk build/tc_harness 35
done
