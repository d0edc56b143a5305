import sys; sys.path.insert(0,'.')
import numpy as np, torch
from adaptive_optics_gym_b200 import AOVecEnv
from oracle.ao_oracle import hcipy_make_pupil_grid, hcipy_Cn_squared_from_fried_parameter, von_karman_screen
def scr(seed, r0=0.12):
    g = hcipy_make_pupil_grid(240, 0.5); cn2 = hcipy_Cn_squared_from_fried_parameter(r0, 2.2e-6)
    return von_karman_screen(g, cn2, 10.0, np.random.default_rng(seed))
B=9
for n in (6,7,8):
    kw = dict(atm_fried=0.12, act_type='num_actuators', act_dim=64, obs_dim=n, rew_type='smf_ssim', timesteps_per_episode=3)
    s = np.stack([scr(90+i) for i in range(B)])
    envs = {p: AOVecEnv(B, **kw, initial_screens=s, precision=p) for p in ('f64','tensor','fused')}
    for e in envs.values(): e.reset()
    a = torch.from_numpy(np.random.default_rng(n).normal(0,0.7,(B,64)).astype(np.float32)).cuda()
    for e in envs.values(): e.step(a)
    torch.cuda.synchronize()
    ref = envs['f64'].obs_f64.cpu().numpy()
    for p in ('tensor','fused'):
        o = envs[p].obs_f64.cpu().numpy()
        rel = np.abs(o-ref)/np.abs(ref)
        i = np.unravel_index(np.argmax(rel), rel.shape)
        print(n, p, 'max rel', rel.max(), 'at value/max', ref[i]/ref[i[0]].max(), 'err/max', np.abs(o-ref)[i]/ref[i[0]].max(), 'count>1e-5', int((rel>1e-5).sum()), 'of', rel.size)
    for e in envs.values(): e.close()
