import numpy as np, json, sys
sys.path.insert(0,'.')
from oracle.ao_oracle import OracleAOEnv
from adaptive_optics_gym_b200 import AOEnv
name=sys.argv[1]
z=np.load(f'tests/golden/{name}.npz'); case=json.loads(str(z['case'])); kw=case['kw']
from tools.make_golden import oracle_env, load_ar_tables
ref=oracle_env(case, z["screen"])
lay=ref.layer
tabs=load_ar_tables()
if kw.get('SH_operation'):
    sh=ref.shwfs; idx=sh.estimation_subapertures
    tabs.update(sh_recon=ref.reconstruction_matrix, sh_offset=np.array((sh.mla_x[idx], sh.mla_y[idx]))+ref.slopes_ref)
env=AOEnv(**kw, initial_screen=z['screen'], precision='f64', tables=tabs)
env.reset(); ref.reset()
print('reset', np.max(np.abs(env.last_obs_f64-ref.last_obs_f64)/ref.last_obs_f64))
pos=0
for i in range(2):
    cnt=int(z['noise_counts'][i])*240
    nz=z['noise'][pos:pos+cnt]; pos+=cnt
    if case['action']=='sh':
        a=env.SH_step(noise='none')[0]; ra=np.array(ref.SH_step(poisson=False)[0])
        print('sh action rel', np.max(np.abs(a-ra))/np.max(np.abs(ra)), 'vs golden', np.max(np.abs(a-z['actions'][i]))/np.max(np.abs(ra)))
    else:
        a=ra=z['actions'][i]
    env.step(a, extrusion_noise=nz); ref.step(ra, extrusion_noise=nz.reshape(-1,240))
    s=env._h.get_field('screen'); rs=lay.achromatic_screen
    print(i,'screen max abs diff', np.abs(s-rs).max(), 'scale', np.abs(rs).max())
    print(i,'act diff', np.abs(env._h.get_field('actuators')-ref.deformable_mirror.actuators).max(), np.abs(ref.deformable_mirror.actuators).max())
    print(i,'obs rel vs ref', np.max(np.abs(env.last_obs_f64-ref.last_obs_f64)/ref.last_obs_f64), 'vs golden', np.max(np.abs(env.last_obs_f64-z['obs'][i])/z['obs'][i]), 'ref vs golden', np.max(np.abs(ref.last_obs_f64-z['obs'][i])/z['obs'][i]))
