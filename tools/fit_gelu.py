import numpy as np
from scipy.special import erf
from scipy.optimize import minimize
x = np.linspace(-12, 12, 600001)
g = 0.5*x*(1+erf(x/np.sqrt(2)))
CL = 36.0
def model(c, x):
    x2 = np.minimum(x*x, CL)
    p = c[-1]
    for k in range(len(c)-2, -1, -1):
        p = p*x2 + c[k]
    u = x*p   # gelu = x * sigmoid(2u) = x / (1 + exp(-2u))
    return x/(1+np.exp(-2*u))
def maxerr(c): return np.abs(model(c,x)-g).max()
for c0 in ([0.797507856656817, 0.037005669168905415, -0.0003515202679386988],
           [0.7976056508137875, 0.03686192166680285, -0.00030257932964505884, -4.213793400646869e-06]):
    best=np.array(c0)
    for rounds in range(6):
        r = minimize(maxerr, best, method='Nelder-Mead', options={'xatol':1e-12,'fatol':1e-12,'maxiter':4000})
        best=r.x
    e=np.abs(model(best,x)-g)
    print(len(c0), best.tolist(), e.max(), x[e.argmax()])
    # relative-ish error where |g|>1e-3
    m=np.abs(g)>1e-2
    print("  max rel err (|g|>1e-2):", (e[m]/np.abs(g[m])).max())
