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
        # collected time-major (every step writes one contiguous [B, ...] slab), transposed once at the end
        obs_b = torch.empty((E, T, B, n2), **kw)
        act_b = torch.empty((E, T, B, K), **kw)
        logp_b = torch.empty((E, T, B), **kw)
        rew_b = torch.empty((E, T, B), **kw)
        next_b = torch.empty((E, T, B, n2), **kw)
        done_b = torch.zeros((E, T, B), **kw)
        for e in range(E):
            obs, _ = env.reset()
            for t in range(T):
                o32 = obs_b[e, t]
                o32.copy_(obs)                                  # float16 -> float32 straight into the batch
                if self.policy is None:
                    action, logp = env.SH_step()
                    logp = logp.to(torch.float32)
                else:
                    with torch.no_grad():
                        action, logp = self.policy(o32)
                    if self.action_noise is not None:
                        action = action + self.action_noise(action)
                obs, rew, done, _, _ = env.step(action)
                act_b[e, t].copy_(action)
                logp_b[e, t].copy_(logp)
                rew_b[e, t].copy_(rew)
                next_b[e, t].copy_(obs)
                done_b[e, t].copy_(done)
            self.num_episodes += B
        N = E * B * T

        def rows(x):                                            # [E, T, B, ...] -> (episode, env, step) rows
            return x.transpose(1, 2).reshape((N,) + tuple(x.shape[3:]))

        rew_rows = rows(rew_b)
        self.batch_ep_rew = rew_rows.reshape(E * B, T)
        lens = torch.full((E * B,), float(T), **kw)          # every env terminates on step T (AO_env.py:147)
        return rows(obs_b), rows(act_b), rows(logp_b), rew_rows, rows(next_b), rows(done_b), lens

    def episode_returns(self):
        """undiscounted return of every episode of the last ``rollout`` ([episodes] float32 cuda)"""
        return self.batch_ep_rew.sum(dim=1)
