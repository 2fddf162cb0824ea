"""Configuration object of the training path.

Mirror of the reference ``arguments.py:74-106`` (the hard-coded ``Args`` class is the reference's real
config; its argparse ``get_args`` is dead code and is not reproduced).  Field names and default values
are the reference's; the fields below the marker are additions for the vectorised GPU path.
"""


class Args:
    def __init__(self):
        self.n_epochs = 200
        self.n_cycles = 50
        self.n_batches = 40
        self.save_interval = 5
        self.seed = 125
        self.num_workers = 19          # unused by the reference as well
        self.replay_strategy = 'future'
        self.clip_return = 50          # unused by the reference: the clamp is 1/(1-gamma) (ddpg_agent.py:259)
        self.save_dir = 'saved_models/'
        self.noise_eps = 0.01
        self.random_eps = 0.3
        self.buffer_size = 1e6 * 1 / 2
        self.replay_k = 4
        self.clip_obs = 200
        self.batch_size = 256
        self.gamma = 0.98
        self.action_l2 = 1
        self.lr_actor = 0.001
        self.lr_critic = 0.001
        self.polyak = 0.95
        self.n_test_rollouts = 25
        self.clip_range = 5
        self.demo_length = 25
        self.cuda = True               # the reference default is False (CPU torch); this path is CUDA-only
        self.num_rollouts_per_mpi = 2
        self.add_demo = True
        self.demo_name = "bmirobot_1000_push_demo.npz"
        self.train_type = "push"       # or "pick"
        self.Use_GUI = False           # reference default True opens an OpenGL window (arguments.py:105)
        self.env_name = 'bmirobot_' + str(self.train_type) + " seed" + str(self.seed)
        # ---- additions for the vectorised B200 path ------------------------------------------------
        self.n_envs = None             # envs stepped per kernel launch; None -> num_rollouts_per_mpi
        self.buffer_dtype = "float32"  # "float64" reproduces the reference's storage bit for bit
        self.device_rng = True         # Philox draws inside the captured graphs; False -> numpy global stream
        self.late_clip_epoch = 100     # epoch >= 100 clips rollout actions to +-0.15 (ddpg_agent.py:118-119)
        self.use_cuda_graphs = True
        self.p2p_adam = True           # multi-GPU: fused peer-memory gradient sum + Adam (False: NCCL allreduce, then Adam)
        self.fused_rollout = True      # one kernel launch per batch of episodes (policy MLP inside the env kernel)
        self.verbose = True
