timeout 200 python tools_dev/ab_1x1.py
