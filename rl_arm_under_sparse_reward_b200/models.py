"""Actor / critic containers.

Mirror of the reference ``models.py:11-44``: ``actor(env_params)`` 30->256->256->256->4 with
``max_action * tanh`` and ``critic(env_params)`` 34->256->256->256->1 on ``cat[x, a/max_action]``;
``state_dict`` keys ``fc1 fc2 fc3 action_out|q_out`` so checkpoints written by the agent load in
the reference's ``demo_push.py:40-41``.  All parameters alias ONE flat float32 buffer in
named_parameters order (utils.py:18-27) — the layout the CUDA learner (csrc/ddpg.cu), the
NCCL broadcast/allreduce and the Adam kernel work on.  ``forward`` is provided for API
compatibility and tests; the training / rollout hot path calls ``bmi_ddpg_*`` on the flat
buffers instead.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def _flatten_parameters_(module, device):
    params = list(module.parameters())
    flat = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device=device)
    off = 0
    for p in params:
        n = p.numel()
        flat[off:off + n].copy_(p.data.reshape(-1))
        p.data = flat[off:off + n].view(p.shape)
        off += n
    module.flat = flat
    module.flat_grad = None  # set by the trainer to its slice of the flat gradient buffer
    return flat


class actor(nn.Module):
    def __init__(self, env_params, device=None):
        super().__init__()
        self.max_action = env_params['action_max']
        n_in = env_params['obs'] + env_params['goal']
        self.fc1 = nn.Linear(n_in, 256)
        self.fc2 = nn.Linear(256, 256)
        self.fc3 = nn.Linear(256, 256)
        self.action_out = nn.Linear(256, env_params['action'])
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu"
        _flatten_parameters_(self, device)

    def forward(self, x):
        h = F.relu(self.fc1(x))
        h = F.relu(self.fc2(h))
        h = F.relu(self.fc3(h))
        return self.max_action * torch.tanh(self.action_out(h))


class critic(nn.Module):
    def __init__(self, env_params, device=None):
        super().__init__()
        self.max_action = env_params['action_max']
        n_in = env_params['obs'] + env_params['goal'] + env_params['action']
        self.fc1 = nn.Linear(n_in, 256)
        self.fc2 = nn.Linear(256, 256)
        self.fc3 = nn.Linear(256, 256)
        self.q_out = nn.Linear(256, 1)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cpu"
        _flatten_parameters_(self, device)

    def forward(self, x, actions):
        h = torch.cat([x, actions / self.max_action], dim=1)
        h = F.relu(self.fc1(h))
        h = F.relu(self.fc2(h))
        h = F.relu(self.fc3(h))
        return self.q_out(h)
