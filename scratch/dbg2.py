import numpy as np, sys
sys.path.insert(0,'.')
from mcos_b200.engine import Engine
n,t,nsim=130,520,2
rng = np.random.RandomState(n + t)
a = rng.standard_normal((n, 3 * n))
cov = a @ a.T / (3 * n) * np.outer(*(2 * [rng.uniform(0.05, 0.2, n)]))
mu = np.zeros(n)
covs = np.stack([np.cov(rng.multivariate_normal(mu, cov, size=t), rowvar=False).reshape(n, n) for _ in range(nsim)])
eng=Engine(n,t)
out,info=eng.denoise(covs[1:2], t)
print(info['var'], info['n_facts'], info['status'])
out,info=eng.denoise(covs, t)
print(info['var'], info['n_facts'], info['status'])
