for v in "" _epv1 _epv2 _epv3; do
export FOCR_B200_LIB=$PWD/font-ocr_b200/libfocr_b200$v.so
echo "variant $v"; timeout 300 python tools/prof_run.py 16 0.8 2>&1 | grep "scan\|Error" | tail -2
done
