for bq in 32 8 4; do for occ in 2 4; do
echo "== batch=$bq occ=$occ"
PCS_ICP_RINGS=1 PCS_ICP_MODE=0 PCS_ICP_BATCH=$bq PCS_ICP_OCC=$occ PCS_TRACK_TIMING=1 python tools/time_tracking.py 198 2 2>&1 | grep -E "icp level|track timing" | tail -4
done; done
