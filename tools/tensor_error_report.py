import numpy as np, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptive_optics_gym_b200 import AOEnv
from oracle.ao_oracle import OracleAOEnv
from tests.test_parity_gpu import _screen
def run(kw, K, nsteps=6, seed=0, r0=0.15):
    scr=_screen(seed, r0)
    env=AOEnv(precision='tensor', **kw, initial_screen=scr); ref=OracleAOEnv(**kw, initial_screen=scr)
    env.reset(); ref.reset()
    rng=np.random.default_rng(1); eo=ep=er=es=0
    e0=np.max(np.abs(env.last_obs_f64-ref.last_obs_f64)/ref.last_obs_f64)
    for t in range(nsteps):
        a=rng.uniform(-1,1,K).astype(np.float32)
        o,r,d,_,i=env.step(a); ro,rr,rd,_,ri=ref.step(a)
        eo=max(eo,np.max(np.abs(env.last_obs_f64-ref.last_obs_f64)/ref.last_obs_f64))
        ep=max(ep,abs(i['power']-ri['power'])/ri['power']); er=max(er,abs(r-rr)/abs(rr))
        if kw.get('rew_type','strehl_ratio')=='strehl_ratio': es=max(es,abs(env.last_strehl-ref.last_strehl)/ref.last_strehl)
    print(kw.get('act_type','num_actuators'),K,'reset obs',f'{e0:.2e}','obs',f'{eo:.2e}','power',f'{ep:.2e}','reward',f'{er:.2e}','strehl',f'{es:.2e}')
    env.close()
for s in range(3):
    run(dict(atm_fried=0.2,act_dim=64,obs_dim=2,rew_type='strehl_ratio',timesteps_per_episode=50),64,seed=s,r0=0.2)
    run(dict(act_type='zernike',act_dim=6,obs_dim=5,rew_type='smf_ssim',timesteps_per_episode=50),6,seed=10+s)
    run(dict(atm_fried=0.08,act_dim=64,obs_dim=5,rew_type='strehl_ratio',timesteps_per_episode=50),64,seed=20+s,r0=0.08)
