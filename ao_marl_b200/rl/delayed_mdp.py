"""Delayed credit assignment for a loop with command latency.

Behavioural twin of the reference's ``DelayedMDP``
(src/reinforcement_learning/environment/delayed_mdp.py:5-58, used by train_rpc.py:530-535, 620-631,
734-757): transitions are held in a window of ``delay + int(not modification)`` steps; once the
window is full, the OLDEST (state, action) is paired with the NEWEST next-state and with the reward
observed at the current step.  With controller delay 1 this yields (s, a, s'', r''') and with delay
0 (s, a, s', r''), as the reference's docstring (train_rpc.py:620-630) requires.

One window of records instead of five parallel deques; entries are stored by reference, so batched
device tensors [E, ...] (or per-agent dicts of them) cost nothing to buffer.
"""
from collections import deque, namedtuple

_Record = namedtuple("_Record", "state action next_state next_action")


class DelayedMDP:
    def __init__(self, delay, modification):
        self.delay = int(delay)
        self.not_modification = int(not modification)
        self.depth = self.delay + self.not_modification
        self._window = deque(maxlen=self.depth)
        self._rewards = deque(maxlen=self.depth)

    # views with the reference's attribute names (read-only)
    @property
    def state_list(self):
        return [r.state for r in self._window]

    @property
    def action_list(self):
        return [r.action for r in self._window]

    @property
    def next_state_list(self):
        return [r.next_state for r in self._window]

    @property
    def next_action_list(self):
        return [r.next_action for r in self._window]

    @property
    def reward_list(self):
        return list(self._rewards)

    def check_update_possibility(self):
        """True once enough steps are buffered for the delay to be accounted for."""
        return len(self._window) >= self.depth

    def save(self, s, a, s_next, r=None):
        self._window.append(_Record(s, a, s_next, None))
        if r is not None:
            self._rewards.append(r)

    def credit_assignment(self):
        return self._window[0].state, self._window[0].action, self._window[-1].next_state

    def save_agent_deactivated(self, s, a, s_next, a_next):
        self._window.append(_Record(s, a, s_next, a_next))

    def credit_assignment_agent_deactivated(self):
        first = self._window[0]
        return first.state, first.action, first.next_state, first.next_action

    def obtain_rewards(self):
        return sum(self._rewards)

    def model_information(self):
        return self._window[0].state, self._window[0].next_state, self._rewards[-1]
