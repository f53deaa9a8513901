import os, sys
sys.path.insert(0, os.getcwd())
import rl_6_nimmt_b200
from rl_6_nimmt_b200 import _native as N
N.LIB_PATH = os.path.join(os.path.dirname(N.LIB_PATH), "dbg", "libnimmt_b200.so")
sys.argv = ["policy_time.py", "4"]
exec(open("profiles/tools/policy_time.py").read())
