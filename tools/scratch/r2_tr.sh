export FOCR_B200_LIB=$PWD/font-ocr_b200/libfocr_b200_exp.so
timeout 300 python tools/tc_modes.py 0,8,32,40 16 2>&1 | tail -5
export FOCR_B200_LIB=$PWD/font-ocr_b200/libfocr_b200_expbase.so
timeout 300 python tools/tc_modes.py 0,32 16 2>&1 | tail -3
