"""Host-side reference forward of the PMC policy (networks/legged_robot/pmc_net/pmc_net.py:33-58,99-178 with
policy_config of test_scripts/primitive_level/test_primitive_level_env.py:39-57): observation normalisation with the stored
running mean / std clipped to +-5 (layers.py:55, pmc_net.py:130-135), VQ encoder 207 -> 256 -> 256 -> 32 with a 32 x 256
codebook (argmin of the squared distance, pmc_net.py:157-171), low-level controller (prop 135 -> 64) + (z 32 -> 32) -> 256 ->
256 -> 12, ReLU activations; the deterministic action is the mean (agent.step(argmax=True)).

`weights` is the list of 28 arrays of a shipped ``*.model`` file in its stored order:
0-3 prop / future running mean, std | 4-9 value head | 10-15 encoder | 16 codebook | 17-20 bottleneck embeds | 21-26 decoder | 27 logstd
"""
import numpy as np


class PmcPolicy:
    def __init__(self, weights):
        w = [np.asarray(x, dtype=np.float32) for x in weights]
        assert len(w) == 28 and w[10].shape == (207, 256) and w[16].shape == (32, 256) and w[25].shape == (256, 12)
        self.prop_mean, self.prop_std, self.fut_mean, self.fut_std = w[0][0], w[1][0], w[2][0], w[3][0]
        self.enc = [(w[10], w[11]), (w[12], w[13]), (w[14], w[15])]
        self.codebook = w[16]
        self.prop_embed, self.z_embed = (w[17], w[18]), (w[19], w[20])
        self.dec = [(w[21], w[22]), (w[23], w[24]), (w[25], w[26])]
        self.logstd = w[27]

    def normalise(self, obs):
        p = np.clip((obs[:, :135] - self.prop_mean) / (self.prop_std + 1e-8), -5.0, 5.0)
        f = np.clip((obs[:, 135:207] - self.fut_mean) / (self.fut_std + 1e-8), -5.0, 5.0)
        return p, f

    def encode(self, p, f):
        x = np.concatenate([p, f], axis=1)
        x = np.maximum(x @ self.enc[0][0] + self.enc[0][1], 0.0)
        x = np.maximum(x @ self.enc[1][0] + self.enc[1][1], 0.0)
        z = x @ self.enc[2][0] + self.enc[2][1]
        d = (z ** 2).sum(1, keepdims=True) - 2.0 * z @ self.codebook + (self.codebook ** 2).sum(0, keepdims=True)
        idx = np.argmax(-d, axis=1)
        return z, idx

    def act(self, obs, return_code=False):
        obs = np.asarray(obs, dtype=np.float32)
        p, f = self.normalise(obs)
        z, idx = self.encode(p, f)
        zq = self.codebook.T[idx]
        pe = np.maximum(p @ self.prop_embed[0] + self.prop_embed[1], 0.0)
        ze = np.maximum(zq @ self.z_embed[0] + self.z_embed[1], 0.0)
        x = np.concatenate([pe, ze], axis=1)
        x = np.maximum(x @ self.dec[0][0] + self.dec[0][1], 0.0)
        x = np.maximum(x @ self.dec[1][0] + self.dec[1][1], 0.0)
        a = x @ self.dec[2][0] + self.dec[2][1]
        return (a, idx) if return_code else a
