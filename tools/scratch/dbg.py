import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from conftest import synthetic_pair
import methods.iterative as it
from oracle import reference_numpy as oracle
_, r = synthetic_pair(20, 30, 25)
t = np.full((20, 30, 3), 0.5)
np.random.seed(9)
ref, tro = oracle.idt_instrumented(t, r)
np.random.seed(9)
tr = {}
out = it.iterative_distribution_transfer(t, r, trace=tr)
print('max err', np.abs(out-ref).max())
for i in range(4):
    print(i, 'counts_t eq', np.array_equal(tr['counts_t'][i], tro[i]['counts_t']), 'counts_r eq', np.array_equal(tr['counts_r'][i], tro[i]['counts_r']),
          'lo', tr['lo'][i]-tro[i]['lo'], 'hi', tr['hi'][i]-tro[i]['hi'], 'lut maxdiff', np.abs(tr['lut'][i]-tro[i]['lut']).max())
    print('  state spread oracle', np.ptp(tro[i]['state'],axis=0), 'nz bins t', [np.nonzero(tro[i]['counts_t'][j])[0].tolist() for j in range(3)], [np.nonzero(tr['counts_t'][i][j])[0].tolist() for j in range(3)])
print(out[0,0], ref[0,0])
