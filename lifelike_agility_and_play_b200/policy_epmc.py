"""Host restatement (numpy) of the environmental-level policy's inference path, for evaluation runs on the vectorised engine
(SURVEY row f2 for EPMC; tools/statistical_pin_epmc.py).

Follows networks/legged_robot/epmc_net/epmc_net.py with the shipped actor configuration (train_scripts/example_epmc_train.sh:53-82:
llc_light, discrete_z, expert_lstm, lstm_layer_norm, append_hist_a, rms on the proprioception only):

* `usr_cmd_encoder` (epmc_net.py:120-135): target 3 -> 32 (ReLU) | `percep_2d_encoder` on the 25 x 13 height map (:86-94: 1x1 conv 4,
  4x4 stride 2, 2x2 stride 2, 2x2 -> 1 channel, all SAME + ReLU = the `tf.contrib.layers.conv2d` defaults) -> 28 | `percep_1d_encoder`
  on the 128 lidar rays (:109-117: periodic padding 4, conv1d 4 SAME, crop, two stride-2 convs, one to 1 channel) -> 32 | the same 2-D
  encoder with its own weights on the front map -> 28; concatenated (120) -> 64 (ReLU);
* `mlc_encoder` (:138-166): prop 135 -> 64, concat with the command embedding -> 256 (ReLU) -> LSTM(32, layer norm) -> 256 logits;
  the code index (argmax here, a categorical sample in the actor) selects a column of the primitive-level codebook (`mapping_z`,
  :169-177);
* `llc` (pmc_net.py:99-112): the frozen primitive-level decoder, prop 135 -> 64 | z 32 -> 32 -> 256 -> 256 -> 12.

The LSTM comes from `tpolicies` (TLeague's policy library, absent from the reference tree): restated from its published form --
`z = ln(x wx) + ln(h wh) + b`, gates `i, f, o, u`, `f = sigmoid(f + forget_bias)`, `h = o tanh(ln(c))`, state `[c, h]`, both zeroed
where the mask (episode start) is set -- with the variable order of the shipped files: wx, wh, b, then (beta, gamma) of the three layer
norms (x, h, c), as `tf.contrib.layers.layer_norm` creates them; the three additive vectors carry identical values in the shipped
files (same initialiser, same gradient), the two 128-vectors and the 32-vector with means 2.2 / 1.1 / 3.6 are the gains.  Nothing pins
this restatement bit for bit (no TensorFlow here); it is pinned behaviourally: the shipped Bullet-trained weights have to traverse the
corridors on the engine.

`weights` = the list of 102 arrays of a shipped ``environmental_level_*.model``:
0-1 prop running mean / std | 2-46 value tower (unused here except by `value`) | 47-48 prop embed | 49-76 command encoder |
77-78 embed | 79-87 LSTM | 88-89 logits | 90 codebook | 91-100 low-level controller | 101 logstd
"""
import numpy as np


def _same_pad(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2, total - total // 2


def conv2d_same_relu(x, w, b, stride):
    """x [B, H, W, C], w [kh, kw, C, O] (TF layout), SAME padding, ReLU."""
    B, H, W, C = x.shape
    kh, kw, _, O = w.shape
    oh, pt, pb = _same_pad(H, kh, stride)
    ow, pl, pr = _same_pad(W, kw, stride)
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    out = np.zeros((B, oh, ow, O), np.float32)
    for i in range(kh):
        for j in range(kw):
            patch = xp[:, i:i + (oh - 1) * stride + 1:stride, j:j + (ow - 1) * stride + 1:stride, :]
            out += patch @ w[i, j]
    return np.maximum(out + b, 0.0)


def conv1d_same_relu(x, w, b, stride):
    """x [B, W, C], w [k, C, O], SAME padding, ReLU."""
    return conv2d_same_relu(x[:, None], w[None], b, stride)[:, 0] if stride == 1 else \
        _conv1d_strided(x, w, b, stride)


def _conv1d_strided(x, w, b, stride):
    B, W, C = x.shape
    k, _, O = w.shape
    ow, pl, pr = _same_pad(W, k, stride)
    xp = np.pad(x, ((0, 0), (pl, pr), (0, 0)))
    out = np.zeros((B, ow, O), np.float32)
    for j in range(k):
        out += xp[:, j:j + (ow - 1) * stride + 1:stride, :] @ w[j]
    return np.maximum(out + b, 0.0)


def _ln(x, g, b, eps=1e-12):        # tf.contrib.layers.layer_norm: last axis, variance_epsilon 1e-12
    m = x.mean(1, keepdims=True)
    v = ((x - m) ** 2).mean(1, keepdims=True)
    return (x - m) / np.sqrt(v + eps) * g + b


def _sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


class LnLstm:
    def __init__(self, w, forget_bias=1.0):
        self.wx, self.wh, self.b, self.bx, self.gx, self.bh, self.gh, self.bc, self.gc = w
        self.nh = self.wh.shape[0]
        self.forget_bias = forget_bias

    def step(self, x, state, mask):
        """x [B, nin], state [B, 2 nh] = [c, h], mask [B] (1 = the episode starts with this step)."""
        keep = (1.0 - mask)[:, None]
        c, h = state[:, :self.nh] * keep, state[:, self.nh:] * keep
        z = _ln(x @ self.wx, self.gx, self.bx) + _ln(h @ self.wh, self.gh, self.bh) + self.b
        i, f, o, u = np.split(z, 4, axis=1)
        c = _sigmoid(f + self.forget_bias) * c + _sigmoid(i) * np.tanh(u)
        h = _sigmoid(o) * np.tanh(_ln(c, self.gc, self.bc))
        return h, np.concatenate([c, h], axis=1).astype(np.float32)


class CommandEncoder:
    def __init__(self, w):
        """w = the arrays of one `usr_cmd_encoder` in creation order: percep_2d (8), percep_1d (8), percep_front (8), [vector feature
        (2),] fc (2) -- 28 with the vector feature (the target direction), 26 without (sepmc_net.py:149-173 before `target_info` exists)."""
        self.c2d, self.c1d, self.cfr = w[0:8], w[8:16], w[16:24]
        if len(w) == 28:
            self.wt, self.bt, self.wf, self.bf = w[24], w[25], w[26], w[27]
        else:
            assert len(w) == 26
            self.wt, self.bt, self.wf, self.bf = None, None, w[24], w[25]

    @staticmethod
    def _enc2d(x, w):
        e = x[..., None].astype(np.float32)
        e = conv2d_same_relu(e, w[0], w[1], 1)
        e = conv2d_same_relu(e, w[2], w[3], 2)
        e = conv2d_same_relu(e, w[4], w[5], 2)
        e = conv2d_same_relu(e, w[6], w[7], 1)
        return e.reshape(e.shape[0], -1)

    @staticmethod
    def _enc1d(x, w, k=4):
        p = np.concatenate([x[:, -k:], x, x[:, :k]], axis=1)[..., None].astype(np.float32)     # periodic padding (epmc_net.py:97-106)
        e = conv1d_same_relu(p, w[0], w[1], 1)[:, k:-k, :]
        e = conv1d_same_relu(e, w[2], w[3], 2)
        e = conv1d_same_relu(e, w[4], w[5], 2)
        e = conv1d_same_relu(e, w[6], w[7], 1)
        return e.reshape(e.shape[0], -1)

    def __call__(self, percep_2d, percep_1d, percep_front, target=None):
        parts = [self._enc2d(percep_2d, self.c2d), self._enc1d(percep_1d, self.c1d), self._enc2d(percep_front, self.cfr)]
        if self.wt is not None:
            parts = [np.maximum(target @ self.wt + self.bt, 0.0)] + parts
        e = np.concatenate(parts, axis=1)
        return np.maximum(e @ self.wf + self.bf, 0.0)


class EpmcPolicy:
    """Deterministic inference (argmax code, mean action) of the shipped environmental-level policy on [N, 916] engine observations."""

    def __init__(self, weights):
        w = [np.asarray(a, np.float32) for a in weights]
        assert len(w) == 102 and w[0].shape == (1, 135) and w[90].shape == (32, 256), "not an environmental-level model"
        self.mean, self.std = w[0], w[1]
        self.vf_fc1, self.vf_cmd, self.vf_fc2, self.vf_fc3 = (w[2], w[3]), CommandEncoder(w[4:32]), (w[32], w[33]), (w[34], w[35])
        self.vf_lstm, self.vf_out = LnLstm(w[36:45]), (w[45], w[46])
        self.prop_embed, self.cmd, self.embed = (w[47], w[48]), CommandEncoder(w[49:77]), (w[77], w[78])
        self.lstm, self.logits = LnLstm(w[79:88]), (w[88], w[89])
        self.codebook = w[90]
        self.llc_prop, self.llc_z = (w[91], w[92]), (w[93], w[94])
        self.dec = [(w[95], w[96]), (w[97], w[98]), (w[99], w[100])]
        self.logstd = w[101]
        self.nh = 32

    def initial_state(self, n):
        return np.zeros((n, 2 * self.nh), np.float32)

    @staticmethod
    def split(obs):
        o = np.asarray(obs, np.float32)
        return (o[:, 0:135], o[:, 135:460].reshape(-1, 25, 13), o[:, 460:588], o[:, 588:913].reshape(-1, 25, 13), o[:, 913:916])

    def act(self, obs, state, mask, rng=None, return_code=False):
        """obs [N, 916] (prop 99 | prop_a 36 | percep_2d 325 | percep_1d 128 | percep_front 325 | target 3), state [N, 64] of the z-LSTM,
        mask [N] = 1 where the observation is the first of an episode.  `rng`: sample the code from the logits (the actor's
        behaviour) instead of the argmax.  Returns (action [N, 12], new state)."""
        prop, p2d, p1d, pfr, tgt = self.split(obs)
        p = np.clip((prop - self.mean) / (self.std + 1e-8), -5.0, 5.0)
        pe = np.maximum(p @ self.prop_embed[0] + self.prop_embed[1], 0.0)
        ce = self.cmd(p2d, p1d, pfr, tgt)
        e = np.maximum(np.concatenate([pe, ce], axis=1) @ self.embed[0] + self.embed[1], 0.0)
        h, state = self.lstm.step(e, state, np.asarray(mask, np.float32))
        logits = h @ self.logits[0] + self.logits[1]
        if rng is None:
            code = logits.argmax(1)
        else:
            g = -np.log(-np.log(rng.uniform(1e-12, 1.0, logits.shape)))
            code = (logits + g).argmax(1)
        z = self.codebook.T[code]
        a = np.maximum(p @ self.llc_prop[0] + self.llc_prop[1], 0.0)
        b = np.maximum(z @ self.llc_z[0] + self.llc_z[1], 0.0)
        x = np.concatenate([a, b], axis=1)
        x = np.maximum(x @ self.dec[0][0] + self.dec[0][1], 0.0)
        x = np.maximum(x @ self.dec[1][0] + self.dec[1][1], 0.0)
        act = (x @ self.dec[2][0] + self.dec[2][1]).astype(np.float32)
        return (act, state, code) if return_code else (act, state)


class SepmcPolicy:
    """Deterministic inference of the shipped strategic-level policy (networks/legged_robot/sepmc_net/sepmc_net.py, actor configuration
    of test_scripts/strategic_level/test_strategic_level_env.py:44-75) on [N, 965] engine observations: `hlc_encoder` (sepmc_net.py:122-146:
    prop 135 -> 64 | perception 88 -> 64 | game vector 29 -> 64 -> 64, concat 192 -> 256 -> LSTM(32) -> heading angle, clipped to +-pi),
    whose (cos, sin) joins the commanded speed as the `target_info` of the environmental-level encoder (`mlc_encoder`, :176-203) -> 256-way
    code -> the frozen primitive-level decoder.  `weights` = the 152 arrays of ``strategic_level.model``:
    0-1 rms | 2-50 value tower | 51-96 heading controller | 97-139 code controller | 140 codebook | 141-150 decoder | 151 logstd."""

    def __init__(self, weights):
        w = [np.asarray(a, np.float32) for a in weights]
        assert len(w) == 152 and w[83].shape == (192, 256) and w[140].shape == (32, 256), "not a strategic-level model"
        self.mean, self.std = w[0], w[1]
        self.h_prop, self.h_percept = (w[51], w[52]), CommandEncoder(w[53:79])
        self.h_vec = [(w[79], w[80]), (w[81], w[82])]
        self.h_embed, self.h_lstm, self.h_mu = (w[83], w[84]), LnLstm(w[85:94]), (w[94], w[95])
        self.m_prop, self.m_cmd, self.m_embed = (w[97], w[98]), CommandEncoder(w[99:127]), (w[127], w[128])
        self.m_lstm, self.logits = LnLstm(w[129:138]), (w[138], w[139])
        self.codebook = w[140]
        self.llc_prop, self.llc_z = (w[141], w[142]), (w[143], w[144])
        self.dec = [(w[145], w[146]), (w[147], w[148]), (w[149], w[150])]
        self.nh = 32

    def initial_state(self, n):
        return np.zeros((n, 4 * self.nh), np.float32)          # [c, h] of the heading LSTM, then of the code LSTM

    def act(self, obs, state, mask, return_aux=False):
        o = np.asarray(obs, np.float32)
        prop, p2d, p1d, pfr = o[:, 0:135], o[:, 135:460].reshape(-1, 25, 13), o[:, 460:588], o[:, 588:913].reshape(-1, 25, 13)
        game = np.concatenate([o[:, 913:918], o[:, 918:933], o[:, 948:955], o[:, 962:964]], axis=1)      # percept_vec | oppo_info | flag_info | with_flag
        spd = o[:, 964:965]
        mask = np.asarray(mask, np.float32)
        p = np.clip((prop - self.mean) / (self.std + 1e-8), -5.0, 5.0)
        relu = lambda x: np.maximum(x, 0.0)
        ge = relu(relu(game @ self.h_vec[0][0] + self.h_vec[0][1]) @ self.h_vec[1][0] + self.h_vec[1][1])
        e = relu(np.concatenate([relu(p @ self.h_prop[0] + self.h_prop[1]), self.h_percept(p2d, p1d, pfr), ge], axis=1) @ self.h_embed[0] + self.h_embed[1])
        hh, s_h = self.h_lstm.step(e, state[:, :2 * self.nh], mask)
        ang = np.clip(hh @ self.h_mu[0] + self.h_mu[1], -np.pi, np.pi)
        tgt = np.concatenate([np.cos(ang), np.sin(ang), spd], axis=1).astype(np.float32)
        e2 = relu(np.concatenate([relu(p @ self.m_prop[0] + self.m_prop[1]), self.m_cmd(p2d, p1d, pfr, tgt)], axis=1) @ self.m_embed[0] + self.m_embed[1])
        hm, s_m = self.m_lstm.step(e2, state[:, 2 * self.nh:], mask)
        code = (hm @ self.logits[0] + self.logits[1]).argmax(1)
        z = self.codebook.T[code]
        x = np.concatenate([relu(p @ self.llc_prop[0] + self.llc_prop[1]), relu(z @ self.llc_z[0] + self.llc_z[1])], axis=1)
        x = relu(x @ self.dec[0][0] + self.dec[0][1])
        x = relu(x @ self.dec[1][0] + self.dec[1][1])
        act = (x @ self.dec[2][0] + self.dec[2][1]).astype(np.float32)
        new_state = np.concatenate([s_h, s_m], axis=1)
        return (act, new_state, ang[:, 0], code) if return_aux else (act, new_state)


_ENC = [(1, 1, 1, 4), (4,), (4, 4, 4, 4), (4,), (2, 2, 4, 4), (4,), (2, 2, 4, 1), (1,), (4, 1, 4), (4,), (4, 4, 4), (4,), (4, 4, 4), (4,), (4, 4, 1), (1,),
        (1, 1, 1, 4), (4,), (4, 4, 4, 4), (4,), (2, 2, 4, 4), (4,), (2, 2, 4, 1), (1,)]
_ENC28, _ENC26 = _ENC + [(3, 32), (32,), (120, 64), (64,)], _ENC + [(88, 64), (64,)]
_LSTM = [(256, 128), (32, 128), (128,), (128,), (128,), (128,), (128,), (32,), (32,)]
_LLC = [(135, 64), (64,), (32, 32), (32,), (96, 256), (256,), (256, 256), (256,), (256, 12), (12,), (1, 12)]
# array shapes of the shipped files, in their stored order (environmental_level_*.model: 102 arrays, strategic_level.model: 152)
EPMC_SHAPES = ([(1, 135), (1, 135), (135, 128), (128,)] + _ENC28 + [(64, 128), (128,), (256, 256), (256,)] + _LSTM + [(32, 1), (1,)] +
               [(135, 64), (64,)] + _ENC28 + [(128, 256), (256,)] + _LSTM + [(32, 256), (256,), (32, 256)] + _LLC)
SEPMC_SHAPES = ([(1, 135), (1, 135), (135, 128), (128,)] + _ENC26 + [(64, 128), (128,), (29, 64), (64,), (64, 64), (64,), (64, 128), (128,), (384, 256), (256,)] +
                _LSTM + [(32, 1), (1,)] +
                [(135, 64), (64,)] + _ENC26 + [(29, 64), (64,), (64, 64), (64,), (192, 256), (256,)] + _LSTM + [(32, 1), (1,), (1, 1)] +
                [(135, 64), (64,)] + _ENC28 + [(128, 256), (256,)] + _LSTM + [(32, 256), (256,), (32, 256)] + _LLC)


def random_weights(strategic=False, seed=0):
    """Random weights of the shipped architecture (benchmarks, tests): fan-in scaled normals, positive running std."""
    rng = np.random.default_rng(seed)
    shapes = SEPMC_SHAPES if strategic else EPMC_SHAPES
    w = [(rng.standard_normal(s) / np.sqrt(max(1, int(np.prod(s[:-1]))))).astype(np.float32) for s in shapes]
    w[1] = np.abs(w[1]) + 0.2
    return w


# ---------------------------------------------------------------------------------------------------------------- device side
def hier_role_arrays(strategic):
    """Index (in the shipped file's array list) of the array that plays each role of include/llq_policy.h."""
    if not strategic:
        return [0, 1, 47, 48] + list(range(49, 77)) + [77, 78] + list(range(79, 88)) + [88, 89, 90] + list(range(91, 101))
    mlc = [0, 1, 97, 98] + list(range(99, 127)) + [127, 128] + list(range(129, 138)) + [138, 139, 140] + list(range(141, 151))
    hlc = [51, 52] + list(range(53, 79)) + [79, 80, 81, 82] + [83, 84] + list(range(85, 94)) + [94, 95]
    return mlc + hlc


class DeviceHierPolicy:
    """The environmental- / strategic-level policy on the GPU (csrc/llq_policy_hier.cu through include/llq_policy.h): reads the engine's
    observation rows in place, keeps the LSTM states on the device, writes the actions the fused env step consumes."""

    def __init__(self, weights, device=0):
        import ctypes as C
        from .policy import POLICY_LIB_PATH
        import os
        if not os.path.exists(POLICY_LIB_PATH):
            raise RuntimeError("%s is not built (python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback" % POLICY_LIB_PATH)
        self._C, self.lib = C, C.CDLL(POLICY_LIB_PATH)
        w = [np.ascontiguousarray(a, np.float32).reshape(-1) for a in weights]
        self.strategic = len(w) == 152
        assert len(w) in (102, 152), "expected an environmental-level (102 arrays) or a strategic-level (152 arrays) model"
        w = [np.concatenate([a, np.zeros((-a.size) % 4, np.float32)]) for a in w]        # every array starts on a 16-byte boundary (float4 loads)
        starts = np.concatenate([[0], np.cumsum([a.size for a in w])]).astype(np.int64)
        blob = np.concatenate(w)
        off = np.array([starts[i] for i in hier_role_arrays(self.strategic)], np.int32)
        self.lib.llq_hier_policy_last_error.restype = C.c_char_p
        h = C.c_void_p()
        rc = self.lib.llq_hier_policy_create(blob.ctypes.data_as(C.c_void_p), C.c_int64(blob.size), off.ctypes.data_as(C.c_void_p), C.c_int32(off.size),
                                             C.c_int32(int(self.strategic)), C.c_int32(device), C.byref(h))
        if rc:
            raise RuntimeError("llq_hier_policy_create: %s" % self.lib.llq_hier_policy_last_error().decode())
        self._h = h
        self.state_dim = 128 if self.strategic else 64
        self.obs_dim = 965 if self.strategic else 916

    def forward(self, obs_ptr, obs_ld, n, done_ptr, state_ptr, act_ptr, codes_ptr=None, heading_ptr=None, stream=None):
        C = self._C
        rc = self.lib.llq_hier_policy_forward(self._h, C.c_void_p(obs_ptr), C.c_int64(obs_ld), C.c_int32(n), C.c_void_p(done_ptr or 0), C.c_void_p(state_ptr),
                                              C.c_void_p(act_ptr), C.c_void_p(codes_ptr or 0), C.c_void_p(heading_ptr or 0), C.c_void_p(stream or 0))
        if rc:
            raise RuntimeError("llq_hier_policy_forward: %s" % self.lib.llq_hier_policy_last_error().decode())

    def close(self):
        if self._h:
            self.lib.llq_hier_policy_destroy(self._h)
            self._h = None
