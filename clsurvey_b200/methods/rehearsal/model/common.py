"""Mirror of src/methods/rehearsal/model/common.py:14-118 (RehearsalMemory, compute_offsets) -- host bookkeeping.

The exemplar *keys* (the reference stores image paths) are kept exactly like the reference, bit-exact ring-buffer
semantics; the exemplar *pixels* are additionally cached on the device so that a memory pass never goes back to
disk / JPEG decode (the reference re-reads k*n_memories files per training batch, gem.py:233-237)."""
import torch


class RehearsalMemory(object):
    def __init__(self, n_entries, n_memories, input_dim, device=None):
        self.n_entries, self.n_memories, self.input_dim = n_entries, n_memories, tuple(input_dim)
        self.exemplars = {entry: [None] * n_memories for entry in range(n_entries)}
        self.shape = [n_entries, n_memories, *self.input_dim]
        self.pixels = None if device is None else torch.zeros(n_entries, n_memories, *self.input_dim, device=device)

    def __getitem__(self, item):
        if isinstance(item, tuple):
            entry, ex_idx = item
            if isinstance(ex_idx, list):
                return [self.exemplars[entry][x] for x in ex_idx]
            return self.exemplars[entry][ex_idx]
        if isinstance(item, int):
            return self.exemplars[item]
        raise IndexError("getitem with index :{} NOT VALID".format(item))

    def __setitem__(self, key, value):
        if isinstance(key, tuple):
            entry, ex_idx = key
            self.exemplars[entry][ex_idx] = value
        elif isinstance(key, int):
            self.exemplars[key] = value
        self.n_entries = len(self.exemplars)

    def __len__(self):
        return len(self.exemplars)

    def get_exemplar_lengths(self):
        return {task: len(paths) for task, paths in self.exemplars.items()}


def compute_offsets(task_idx, cum_nc_per_task):
    """common.py:106-118: output slice [offset1, offset2) of task `task_idx` in the shared head."""
    offset1 = 0 if task_idx == 0 else int(cum_nc_per_task[task_idx - 1])
    return offset1, int(cum_nc_per_task[task_idx])
