#!/bin/bash
for cfg in "ICADV_TC_PERSIST=0 ICADV_TC_ONE_CTA=1 ICADV_TC_DBG=0" "ICADV_TC_PERSIST=0 ICADV_TC_ONE_CTA=1 ICADV_TC_DBG=24" "ICADV_TC_PERSIST=0 ICADV_TC_DBG=24" "ICADV_TC_PERSIST=0 ICADV_TC_DBG=0"; do
  echo "=== $cfg"
  env $cfg timeout 120 python scripts/tile_timeline.py "conv linear" 2>&1 | tail -2 | cut -c1-420
done
