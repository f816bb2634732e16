#!/bin/bash
for o in "lm_speculate=3" "lm_speculate=4" "lm_speculate=6" "lm_speculate=8"; do NID_OPTS=$o timeout 300 python tools/time_single.py 2>&1 | tr '\n' ' '; echo; done
