import sys, os, tempfile, difflib
sys.path.insert(0, '.')
from oracle import fortran_io as fio
from phaseshifts_b200 import api
muff = 'tests/golden/Re0001/mufftin_Re_slab.d'
d = tempfile.mkdtemp()
g = [os.path.join(d, n) for n in ("g_phasout", "g_dataph", "g_inpdat")]
o = [os.path.join(d, n) for n in ("o_phasout", "o_dataph", "o_inpdat")]
api.phsh_rel_file(muff, *g, real32=True)
fio.phsh_rel_file(muff, *o, real32=True)
for a, b in zip(g, o):
    la, lb = open(a).read().splitlines(), open(b).read().splitlines()
    nd = 0
    print(a, len(la), len(lb))
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            nd += 1
            if nd <= 6:
                print(i, repr(x)); print(i, repr(y))
    print('ndiff', nd)
