#!/bin/bash
# Developer: torch-op profile (device time, input shapes) of the tracker's host-glue stages, one run per stage.
for st in setup results extract_traces; do
  PCS_PROFILE_STAGE=$st python tools/time_tracking.py ${1:-198} 2 2>&1 | grep -A30 "stage profile" | tail -31
done
