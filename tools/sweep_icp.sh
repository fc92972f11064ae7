for bq in 8 16 32; do
echo "== batch=$bq"
PCS_ICP_BATCH=$bq PCS_TRACK_TIMING=1 python tools/time_tracking.py 198 2 2>&1 | grep -E "icp level 2|track timing" | tail -2
done
