"""dev aid: re-run single problems of tools/fuzz_parity.py by seed.  usage: FUZZ_REF=1 python tools/fuzz_seeds.py seed [seed ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import fuzz_parity as F
for sd in sys.argv[1:]:
    sys.argv = ["fuzz_parity.py", sd, "1"]
    F.main()
