import sys, dataclasses
sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import numpy as np
import oraclelib as O, rtm_gpu_b200 as R
from golden_cases import GOLDEN_CASES
from refcase import data_tiny, rel_l2
from test_gpu_parity import prepare, make_engine
case = dataclasses.replace(GOLDEN_CASES["tiny_te_compen"], NT1=150)
v, vmin, vmax, Index, c = prepare(case)
r_u = [24, 34, 40, 55]; r_x = [20, 31, 64, 100]
seis = np.stack([data_tiny(case, 100 * i)[:, :150] for i in range(4)])
p = O.make_params(case, vmin, vmax, contract=1)
ref = [O.migrate_shot(p, v, c, Index, r_u[m], r_x[m], seis[m]) for m in range(4)]
nbad = 0
for it in range(int(sys.argv[1])):
    for B in (1, 4, 3):
        with make_engine(case, v, vmin, vmax, Index, c, max_batch=B) as e:
            u, d, s = e.migrate(r_u, r_x, seis)
        for m in range(4):
            if not (np.array_equal(u[m], ref[m][0]) and np.array_equal(d[m], ref[m][1])):
                bad = np.argwhere(u[m] != ref[m][0])
                nbad += 1
                print('iter', it, 'B', B, 'shot', m, 'nbad', len(bad), 'rel', rel_l2(u[m], ref[m][0]), 'down rel', rel_l2(d[m], ref[m][1]))
print('total mismatches', nbad)
