import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from PIL import Image
from siftmetal_b200 import Engine
from oracle_lib import Oracle
im = np.array(Image.open('tests/golden/butterfly.png').convert('RGBA'))[:, :, [2,1,0,3]].copy()
h, w = im.shape[:2]
eng = Engine(w, h); res = eng.detect_and_describe([im])
ora = Oracle(w, h); ok, oc = ora.detect(im); od, odc = ora.describe()
d = res.descriptors
df = np.abs(d['features'].astype(int) - od['features'].astype(int))
bad = np.nonzero(df.max(1) > 1)[0]
print('bad descriptors', len(bad), 'of', len(d))
k = res.keypoints[d['keypoint']]
for i in bad[:12]:
    print(i, 'oct', k['octave'][i], 'scale', k['scale'][i], 'sub', k['subScale'][i], 'theta', d['theta'][i], 'abs', k['absoluteX'][i], k['absoluteY'][i], 'maxdiff', df[i].max(), 'sumdiff', df[i].sum(), 'nnz', (df[i]>1).sum())
    print('   gpu', d['features'][i][:32].tolist()); print('   ora', od['features'][i][:32].tolist())
print('by octave bad:', np.bincount(k['octave'][bad], minlength=7), 'all:', np.bincount(k['octave'], minlength=7))
print('theta of bad', np.round(d['theta'][bad][:40], 2))
