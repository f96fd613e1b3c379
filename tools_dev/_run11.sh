echo new; timeout 200 python tools_dev/ab_fin.py
echo prev; DS_LIB_PATH=$PWD/build/lib_prev.so timeout 200 python tools_dev/ab_fin.py
