export FOCR_B200_LIB=$PWD/font-ocr_b200/libfocr_b200_expbase.so
for thr in 0.8 2.0; do
echo "=== base thr $thr"
THR=$thr timeout 300 python tools/tc_timeline.py 2>&1 | tail -34
THR=$thr timeout 300 python tools/tc_trace.py 2>&1 | tail -5
done
