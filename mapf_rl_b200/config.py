"""Defaults of the reference's config.py (config.py:1-65) that the hot path reads.

Only the constants consumed by the environment / replay path are mirrored; the Q-network and
launcher settings are out of scope.
"""
# environment (config.py:4-14)
map_length = 20
num_agents = 6
obs_radius = 4
reward_fn = dict(move=-0.075,
                 stay_on_goal=0,
                 stay_off_goal=-0.075,
                 collision=-0.5,
                 finish=3)
obs_shape = (6, 9, 9)

# replay / DQN constants used by the PER + TD path (config.py:24-43, 65)
gamma = 0.99
batch_size = 192
max_steps = 256
bt_steps = 16
local_buffer_size = max_steps
prioritized_replay_alpha = 0.6
prioritized_replay_beta = 0.4
forward_steps = 2
latent_dim = 256
max_comm_agents = 3      # config.py:58, including the agent itself
max_num_agents = 6       # config.py:51 (spelled max_num_agetns there)
max_map_length = 40      # config.py:52
learning_starts = 50000  # config.py:26

# adaptive curriculum start (config.py:49)
init_set = (1, 10)

REWARD_ORDER = ("move", "stay_on_goal", "stay_off_goal", "collision", "finish")
