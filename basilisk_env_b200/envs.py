"""Single-environment gym surface of the reference, backed by the CUDA step.

`leoPowerAttEnv` mirrors /root/reference/basilisk_env/envs/leoPowerAttitudeEnvironment.py:14-216
(same attribute names, (5,1) float64 observations, 4-tuple step return, `info` keys) and
`LEOPowerAttitudeSimulator` mirrors the object it drives
(/root/reference/basilisk_env/simulators/leoPowerAttitudeSimulator.py:67, :535-652):
`run_sim(action) -> (obs(5,1), sim_states, sim_over)`, `.obs`, `.initial_conditions`, `.simTime`,
`close_gracefully()`.  Both are N=1 views of `LeoPowerAttVecEnv`; the arithmetic runs on the GPU."""
import copy

import numpy as np

from . import initial_conditions as _ic
from . import spaces
from .vec_env import LeoPowerAttVecEnv, BskEnvError, DONE_DECAY

__version__ = "0.1.0"


class LEOPowerAttitudeSimulator:
    """One spacecraft simulation == one env of a 1-env CUDA batch.

    `initial_conditions=None` samples from numpy's global legacy RNG in the reference's order
    (initial_conditions.set_ICs + the three discarded wheel-factory draws)."""

    def __init__(self, dynRate, fswRate, step_duration, initial_conditions=None, device=0, **config):
        self.dynRate, self.fswRate, self.step_duration = dynRate, fswRate, step_duration
        self.simTime = 0.0
        if initial_conditions is None:
            self.initial_conditions = _ic.set_ICs()
            sampled = True
        else:
            self.initial_conditions = initial_conditions
            sampled = False
        self.mass = self.initial_conditions.get("mass")
        self.powerDraw = self.initial_conditions.get("powerDraw")
        overrides = _ic.config_overrides(self.initial_conditions)
        overrides.update(config)
        self._vec = LeoPowerAttVecEnv(1, device=device, dynRate=float(dynRate), fswRate=float(fswRate),
                                      step_duration=float(step_duration), **overrides)
        # the reference consumes these draws in set_dynamics -> balancedHR16Triad(useRandom=True) on EVERY
        # construction, also when ICs are passed in (SIM:301)
        _ic.consume_wheel_factory_draws()
        del sampled
        row = _ic.ic_row(self.initial_conditions)[None, :]
        ob = self._vec.reset_ics(row).cpu().numpy()[0]
        # SIM:347-351: un-normalised initial obs (wheel speed norm in RPM: SIM:306,350), eclipse entry 0
        self.obs = np.zeros([5, 1])
        self.obs[0, 0] = np.linalg.norm(np.asarray(self.initial_conditions["sigma_init"], dtype=float))
        self.obs[1, 0] = np.linalg.norm(np.asarray(self.initial_conditions["omega_init"], dtype=float))
        self.obs[2, 0] = np.linalg.norm(np.asarray(self.initial_conditions["wheelSpeeds"], dtype=float))
        self.obs[3, 0] = self.initial_conditions["storedCharge_Init"] / 3600.0
        self._normalised_initial_ob = ob
        self.sim_states = np.zeros([11, 1])
        self.sim_over = False
        self.modeRequest = None
        self._last = None

    # the environment class drives the fused step (physics + gym bookkeeping) through this
    def _step(self, action):
        a = _decode_action(action)
        obs, rew, done, reason = self._vec.step_host(np.array([a], dtype=np.int32))
        self.simTime += self.step_duration
        sim_obs = self._vec.field("sim_obs").cpu().numpy()[:, 0]
        self.obs = sim_obs.reshape(5, 1).copy()
        self.sim_states = []
        self.sim_over = bool(reason[0] & DONE_DECAY)
        self.modeRequest = str(action)
        return obs[0], float(rew[0]), bool(done[0]), int(reason[0])

    def run_sim(self, action):
        """Advance `step_duration` seconds in the requested mode; returns (obs(5,1), [], sim_over)."""
        self._step(action)
        return self.obs, self.sim_states, self.sim_over

    def close_gracefully(self):
        """The reference unloads SPICE kernels here (SIM:646-652); nothing to unload in this build."""
        return

    def close(self):
        self._vec.close()


def _decode_action(action):
    """`modeRequest = str(action)` compared with "0"/"1"/"2" (SIM:543-574): anything else leaves the
    task enables untouched and still advances time (encoded as -1 for the kernel)."""
    s = str(action)
    return int(s) if s in ("0", "1", "2") else -1


_GymEnv = spaces.gym_env_base()


class leoPowerAttEnv(_GymEnv):
    """Simple attitude/orbit control problem: decide when to point at the ground (reward), at the
    Sun (power) or to dump wheel momentum.  gym API of the reference environment."""

    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    spec = None

    def __init__(self, device=0):
        self.__version__ = __version__
        self._device = device
        self.max_length = int(3 * 180)                 # ENV:25
        self.simulator_init = 0
        self.simulator = None
        self.simulator_backup = None
        self.reward_total = 0
        self.mass = 330.0
        self.powerDraw = -5.
        self.wheel_limit = 3000 * _ic.RPM              # ENV:36
        self.power_max = 20.0
        self.step_duration = 180.
        self.reward_mult = 1. / self.max_length
        self.failure_penalty = 1
        self.observation_space = spaces.Box(-1e16, 1e16, shape=(5, 1))
        self.obs = np.zeros([5, ])
        self.action_space = spaces.Discrete(3)
        self.curr_episode = -1
        self.action_episode_memory = []
        self.curr_step = 0
        self.episode_over = False
        self.debug_states = []
        self.sim_over = False

    def seed(self, seed=None):
        """Seeds numpy's global legacy RNG, which is where `reset()` draws the initial conditions.
        (In the reference `seed()` reaches gym's no-op base implementation -- SURVEY quirk Q8;
        here an explicit seed takes effect, `seed()` without argument stays a no-op.)"""
        if seed is not None:
            np.random.seed(seed)
        return [seed]

    def _new_simulator(self, initial_conditions=None):
        if self.simulator is not None:
            self.simulator.close()
        self.simulator = None
        self.simulator = LEOPowerAttitudeSimulator(.1, 1.0, self.step_duration, initial_conditions, device=self._device,
                                                   max_length=self.max_length, wheel_limit_rpm=self.wheel_limit / _ic.RPM,
                                                   power_max=self.power_max, failure_penalty=float(self.failure_penalty))
        self.simulator_init = 1

    def step(self, action):
        """ob (5,1), reward, episode_over, info -- ENV:65-145."""
        if self.simulator_init == 0:
            # ENV:94-96 constructs the simulator with keyword arguments its __init__ does not accept,
            # i.e. step() before reset() raises in the reference as well
            raise BskEnvError("step() called before reset(): the reference raises here too (TypeError at ENV:95)")
        self.action_episode_memory[self.curr_episode].append(action)
        ob, reward, over, reason = self.simulator._step(action)
        self.obs = self.simulator.obs
        self.debug_states = self.simulator.sim_states
        self.sim_over = self.simulator.sim_over
        self.episode_over = over
        step_reward = reward
        # reward_total mirrors ENV:105,113,122: the kernel accumulates the same sums per env
        self.reward_total += step_reward
        ob = np.asarray(ob, dtype=np.float64).reshape(5, 1)
        if self.episode_over:
            info = {'episode': {'r': self.reward_total, 'l': self.curr_step},
                    'full_states': self.debug_states, 'obs': ob, 'done_reason': reason}
            self.simulator.close_gracefully()
        else:
            info = {'full_states': self.debug_states, 'obs': ob}
        self.curr_step += 1
        return ob, step_reward, self.episode_over, info

    def reset(self):
        self.action_episode_memory.append([])
        self.episode_over = False
        self.curr_step = 0
        self.reward_total = 0
        self._new_simulator(None)
        self.seed()
        ob = copy.deepcopy(self.simulator.obs)
        ob[2] = ob[2] / self.wheel_limit
        ob[3] = ob[3] / self.power_max
        return ob

    def reset_init(self):
        self.action_episode_memory.append([])
        self.episode_over = False
        self.curr_step = 0
        self.reward_total = 0
        initial_conditions = self.simulator.initial_conditions
        self._new_simulator(initial_conditions)
        ob = copy.deepcopy(self.simulator.obs)
        ob[2] = ob[2] / self.wheel_limit
        ob[3] = ob[3] / self.power_max
        return ob

    def render(self, mode='human', close=False):
        return

    def _render(self, mode='human', close=False):
        return

    def _get_state(self):
        return self.simulator.obs

    def close(self):
        if self.simulator is not None:
            self.simulator.close()
            self.simulator = None
            self.simulator_init = 0
from .opnav_env import opNavEnv, scenario_OpNav  # noqa: E402,F401  (reference: envs/__init__.py:2)
