import json, numpy as np, sys, warnings
warnings.simplefilter('ignore')
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import lsqfit_b200 as lb
from oracle.fit import nonlinear_fit
from parity_util import TIGHT, exact_minimum
d = json.load(open('tests/golden/nist.json'))
for name in ['lanczos3', 'gauss1', 'mgh10']:
    pr = [p for p in d['problems'] if p['name']==name][0]
    x = np.array(pr['x'])
    fo = nonlinear_fit(pr['form'], x, pr['y'], pr['ysdev'], prior_mean=pr['prior_mean'], prior_cov=pr['prior_sdev'], p0=pr['p0'], tol=TIGHT, x_scale='jac')
    xe, fe, Je, cove = exact_minimum(fo); sd=np.sqrt(np.diag(cove))
    print(name, 'oracle nit', fo.nit, 'gap', np.max(np.abs(fo.pmean-xe)/sd))
    for pol in [0, 1, 4, 10]:
        fd = lb.nonlinear_fit(data=(x, pr['y'], pr['ysdev']), fcn=pr['form'], prior=(pr['prior_mean'], pr['prior_sdev']), p0=pr['p0'], tol=TIGHT, polish=pol)
        g = fd.J.T @ fd.residuals
        print('   polish', pol, 'nit', fd.nit, fd.fitter_results['status'], 'gap', np.max(np.abs(fd.pmean-xe)/sd), 'g*sd', np.max(np.abs(g)*sd))
