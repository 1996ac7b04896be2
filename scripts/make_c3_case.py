"""Write the synthetic C3 inputs and an input file for one pass of the stopping loop: python scripts/make_c3_case.py <dir> [passes]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moquimc_b200 import configs as K, synthetic as S
root = sys.argv[1]
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
K.c3_case(root)
S.write_input(os.path.join(root, "c3.in"), root, os.path.join(root, "out"), ParticlesPerHistory=400.0, StoppingStatistics="true",
              StoppingCriteria=1.0, StatThreshold=0.5, MaxStatPasses=passes)
print(os.path.join(root, "c3.in"))
