"""torch-CPU fp32 restatement of the DDPG update (TEST ORACLE — see oracle/__init__.py).

Follows ddpg_agent.py:225-277 (_update_network), :220-222 (soft update) and models.py:11-44 with
stock torch autograd + torch.optim.Adam, i.e. exactly what the reference executes on one rank.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


class Actor(nn.Module):
    def __init__(self, n_in=30, n_act=4, max_action=0.5):
        super().__init__()
        self.max_action = max_action
        self.fc1, self.fc2, self.fc3 = nn.Linear(n_in, 256), nn.Linear(256, 256), nn.Linear(256, 256)
        self.action_out = nn.Linear(256, n_act)

    def forward(self, x):
        x = F.relu(self.fc3(F.relu(self.fc2(F.relu(self.fc1(x))))))
        return self.max_action * torch.tanh(self.action_out(x))


class Critic(nn.Module):
    def __init__(self, n_in=34, max_action=0.5):
        super().__init__()
        self.max_action = max_action
        self.fc1, self.fc2, self.fc3 = nn.Linear(n_in, 256), nn.Linear(256, 256), nn.Linear(256, 256)
        self.q_out = nn.Linear(256, 1)

    def forward(self, x, a):
        x = torch.cat([x, a / self.max_action], dim=1)
        return self.q_out(F.relu(self.fc3(F.relu(self.fc2(F.relu(self.fc1(x)))))))


def flat_params(net):
    return torch.cat([p.detach().reshape(-1) for _, p in net.named_parameters()])


def flat_grads(net):
    return torch.cat([p.grad.detach().reshape(-1) for _, p in net.named_parameters()])


def load_flat(net, flat):
    off = 0
    for _, p in net.named_parameters():
        n = p.numel()
        p.data.copy_(flat[off:off + n].reshape(p.shape))
        off += n


class Learner:
    def __init__(self, n_obs=27, n_goal=3, n_act=4, max_action=0.5, gamma=0.98, action_l2=1.0, lr=1e-3, polyak=0.95):
        self.actor, self.critic = Actor(n_obs + n_goal, n_act, max_action), Critic(n_obs + n_goal + n_act, max_action)
        self.actor_t, self.critic_t = Actor(n_obs + n_goal, n_act, max_action), Critic(n_obs + n_goal + n_act, max_action)
        self.actor_t.load_state_dict(self.actor.state_dict())
        self.critic_t.load_state_dict(self.critic.state_dict())
        self.oa = torch.optim.Adam(self.actor.parameters(), lr=lr)
        self.oc = torch.optim.Adam(self.critic.parameters(), lr=lr)
        self.gamma, self.l2, self.max_action, self.polyak = gamma, action_l2, max_action, polyak

    def losses_and_grads(self, x, xn, a, r):
        """ddpg_agent.py:250-275 without the optimiser steps; r is (B,1)."""
        with torch.no_grad():
            qn = self.critic_t(xn, self.actor_t(xn))
            y = torch.clamp(r + self.gamma * qn, -1 / (1 - self.gamma), 0)
        critic_loss = (y - self.critic(x, a)).pow(2).mean()
        pi = self.actor(x)
        actor_loss = -self.critic(x, pi).mean() + self.l2 * (pi / self.max_action).pow(2).mean()
        self.oa.zero_grad()
        actor_loss.backward()
        ga = flat_grads(self.actor).clone()
        self.oc.zero_grad()
        critic_loss.backward()
        gc = flat_grads(self.critic).clone()
        return actor_loss.item(), critic_loss.item(), ga, gc

    def update(self, x, xn, a, r):
        la, lc, ga, gc = self.losses_and_grads(x, xn, a, r)
        # grads of the actor were zeroed by critic's zero_grad? no: separate optimisers; restore
        off = 0
        for _, p in self.actor.named_parameters():
            n = p.numel()
            p.grad = ga[off:off + n].reshape(p.shape).clone()
            off += n
        self.oa.step()
        self.oc.step()
        return la, lc

    def soft_update(self):
        """ddpg_agent.py:220-222"""
        for tgt, src in ((self.actor_t, self.actor), (self.critic_t, self.critic)):
            for tp, p in zip(tgt.parameters(), src.parameters()):
                tp.data.copy_((1 - self.polyak) * p.data + self.polyak * tp.data)
