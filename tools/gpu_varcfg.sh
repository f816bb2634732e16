#!/bin/bash
# time one geometry for build variants and option sets: tools/gpu_varcfg.sh "rows cols cell bins" variant[:opts] ...
cfg=$1; shift
for vo in "$@"; do
  v=${vo%%:*}; o=""; [ "$vo" != "$v" ] && o=${vo#*:}
  if [ "$v" = base ]; then unset NID_B200_LIB; else export NID_B200_LIB=$PWD/build/libvar_$v.so; fi
  echo "== $v $o: $(NID_OPTS=$o timeout 300 python tools/time_config.py $cfg 96 20 2>&1 | grep 'path=sorted want_jac=1')"
done
