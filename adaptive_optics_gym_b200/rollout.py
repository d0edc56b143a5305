"""GPU-resident rollout collection for ``AOVecEnv``.

The reference collects a batch with a per-step Python loop that round-trips NumPy
(``algorithm.py:216-296`` ``rollout``): one env, ``episodes_per_iteration`` episodes of
``timesteps_per_episode`` steps, returning
``(batch_obs, batch_act, batch_log_probs, batch_rew, batch_next_obs, batch_done, batch_lens)``.
``VecRolloutCollector.rollout`` returns the same seven fields in the same order for ``num_envs``
lock-stepped environments, as CUDA tensors that never leave the device: every env contributes its own
episodes, so one call yields ``episodes_per_iteration * num_envs`` episodes.  Rows are ordered
episode-major, then env, then step -- each env's episode is a contiguous run of
``timesteps_per_episode`` rows, exactly the layout ``algorithm.py`` builds for its single env, so its
reward-to-go / advantage code (``algorithm.py:300-360``) applies unchanged to the flattened batch.

``policy(obs) -> (action, log_prob)`` mirrors ``network.py:69`` ``get_action`` (obs ``[B, obs_dim]`` float32
cuda in, action ``[B, act_dim]`` and log-probability ``[B]`` out); ``policy=None`` with an env built with
``SH_operation=True`` drives the mirror with the Shack-Hartmann integrator as ``algorithm.py:253`` does.
"""
from __future__ import annotations


class VecRolloutCollector:
    def __init__(self, env, policy=None, action_noise=None):
        import torch
        self._torch = torch
        self.env = env
        self.policy = policy
        self.action_noise = action_noise          # e.g. DDPG's OU noise (algorithm.py:258-259): callable [B, K] -> [B, K]
        if policy is None and not getattr(env, 'SH_operation', False):
            raise ValueError('policy=None needs an env built with SH_operation=True (the SH integrator acts)')
        self.num_episodes = 0

    def rollout(self, episodes_per_iteration=1):
        """-> (batch_obs [N, n^2], batch_act [N, K], batch_log_probs [N], batch_rew [N], batch_next_obs [N, n^2],
        batch_done [N], batch_lens [episodes_per_iteration * B]) with N = episodes_per_iteration * B * T, float32
        CUDA tensors; ``self.batch_ep_rew`` is ``[episodes_per_iteration * B, T]`` (algorithm.py:288 logger field)."""
        torch, env = self._torch, self.env
        B, T = env.num_envs, env.max_steps
        n2, K = env.single_observation_space.shape[0], env.single_action_space.shape[0]
        E = int(episodes_per_iteration)
        kw = dict(dtype=torch.float32, device=env.device)
        obs_b = torch.empty((E, B, T, n2), **kw)
        act_b = torch.empty((E, B, T, K), **kw)
        logp_b = torch.empty((E, B, T), **kw)
        rew_b = torch.empty((E, B, T), **kw)
        next_b = torch.empty((E, B, T, n2), **kw)
        done_b = torch.zeros((E, B, T), **kw)
        for e in range(E):
            obs, _ = env.reset()
            for t in range(T):
                o32 = obs.to(torch.float32)
                obs_b[e, :, t] = o32
                if self.policy is None:
                    action, logp = env.SH_step()
                    logp = logp.to(torch.float32)
                else:
                    with torch.no_grad():
                        action, logp = self.policy(o32)
                    if self.action_noise is not None:
                        action = action + self.action_noise(action)
                obs, rew, done, _, _ = env.step(action)
                act_b[e, :, t] = action.to(torch.float32)
                logp_b[e, :, t] = logp
                rew_b[e, :, t] = rew.to(torch.float32)
                next_b[e, :, t] = obs.to(torch.float32)
                done_b[e, :, t] = done.to(torch.float32)
            self.num_episodes += B
        self.batch_ep_rew = rew_b.reshape(E * B, T)
        lens = torch.full((E * B,), float(T), **kw)          # every env terminates on step T (AO_env.py:147)
        N = E * B * T
        return (obs_b.reshape(N, n2), act_b.reshape(N, K), logp_b.reshape(N), rew_b.reshape(N),
                next_b.reshape(N, n2), done_b.reshape(N), lens)

    def episode_returns(self):
        """undiscounted return of every episode of the last ``rollout`` ([episodes] float32 cuda)"""
        return self.batch_ep_rew.sum(dim=1)
