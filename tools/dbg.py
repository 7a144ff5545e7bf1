import sys; sys.path.insert(0,'/root/repo')
import numpy as np
import psydiff_b200 as pd
from psydiff_b200 import synth
from oracle.oracle import Oracle
model = pd.newPsydiff(**synth.config_c1(N=10, T=50, eps=np.array([30.0, 1e-6])))
pd.compileModel(model)
g = pd.getGradients(model)
ref = Oracle(model).gradients()
for l,a,b,c in zip(g, g.values(), ref["grad_clean"], ref["grad_ref"]): print(l,a,b,c)
one = pd.newPsydiff(**synth.config_c1(N=10, T=50, eps=np.array([30.0])))
pd.compileModel(one)
print(pd.getGradients(one)); print(Oracle(one).gradients()["grad_clean"])
