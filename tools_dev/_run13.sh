for l in a prev a; do echo $l; DS_LIB_PATH=$PWD/build/lib_$l.so timeout 200 python tools_dev/ab_attn.py; done
echo new; timeout 200 python tools_dev/ab_attn.py
