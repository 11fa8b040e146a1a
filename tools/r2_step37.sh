#!/bin/bash
for d in 0 1 2 3; do echo "== debug $d"; FETAL_B200_WGRAD_DEBUG=$d python tools/wgrad_fixed_cost.py 2>&1 | tail -6; done
