"""Where the 1e-5 bound on dark detector pixels comes from (CPU model, no GPU): the fused / tensor kernels carry the
phase in 2^-22 half-turn fixed point and take sin / cos from the SFU (absolute error 2^-21.4); both put an absolute
floor under every detector AMPLITUDE, so the relative error of a pixel grows as 1 / sqrt(pixel / brightest pixel).
Prints the relative error of the darkest 5x5 detector pixel at r0 = 0.08 m with FP64 accumulation for each source."""
import numpy as np, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaptive_optics_gym_b200.tables import AOConfig, build_tables
from oracle.ao_oracle import hcipy_make_pupil_grid, hcipy_Cn_squared_from_fried_parameter, von_karman_screen
cfg=AOConfig(fried_parameter=0.08,num_modes=64,obs_dim=5)
T=build_tables(cfg,rng=np.random.default_rng(0))
g=hcipy_make_pupil_grid(240,0.5)
m1=T['mft_obs_1']; m2=T['mft_obs_2']; ap=T['aperture'].reshape(240,240)
lam=cfg.wavelength_wfs
rng=np.random.default_rng(5)
for s in range(4):
    scr=von_karman_screen(g,hcipy_Cn_squared_from_fried_parameter(0.08,2.2e-6),10.0,np.random.default_rng(20+s)).reshape(240,240)
    a=rng.normal(0,1,64)/(np.arange(64)+10); surf=(T['dm_modes'].T@a).reshape(240,240); surf*=0.1*2.2e-6/surf.std()
    phi=scr/lam+2*surf*2*np.pi/lam
    def obs(ph, sc=None):
        E=ap*np.exp(1j*ph) if sc is None else ap*(sc[0]+1j*sc[1])
        F=m1@E@m2
        return np.abs(F)**2
    ref=obs(phi)
    # (1) phase quantisation 2^-22 half-turns (atmosphere tile) + DM rounding to same grid
    q=np.pi/2**22
    ph_q=np.round(scr/lam/q)*q+np.round(2*surf*2*np.pi/lam/q)*q
    e1=np.abs(obs(ph_q)-ref)/ref
    # (2) MUFU error model: abs error uniform +-2^-21.4 on sin and cos
    err=2**-21.4
    c=np.cos(phi)+rng.uniform(-err,err,phi.shape); sn=np.sin(phi)+rng.uniform(-err,err,phi.shape)
    e2=np.abs(obs(None,(c,sn))-ref)/ref
    # (3) fp32 rounding of sin, cos
    c=np.cos(phi).astype(np.float32).astype(float); sn=np.sin(phi).astype(np.float32).astype(float)
    e3=np.abs(obs(None,(c,sn))-ref)/ref
    print(f'screen {s}: min pixel / max {ref.min()/ref.max():.2e}; rel err max: phase-quant {e1.max():.2e}  mufu {e2.max():.2e}  fp32 {e3.max():.2e}')
