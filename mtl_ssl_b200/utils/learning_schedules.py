"""Host-side learning-rate schedules (a scalar per step, written into the optimizer's device
`hyper` buffer).  Mirrors /root/reference/object_detection/utils/learning_schedules.py:23-103 and
the dispatch of builders/optimizer_builder.py:69-117."""
import math


def exponential_decay_with_burnin(global_step, learning_rate_base, learning_rate_decay_steps,
                                  learning_rate_decay_factor, burnin_learning_rate=0.0, burnin_steps=0):
    if burnin_learning_rate == 0:
        burnin_learning_rate = learning_rate_base
    if global_step < burnin_steps:
        return float(burnin_learning_rate)
    return float(learning_rate_base * learning_rate_decay_factor ** (global_step // learning_rate_decay_steps))


def manual_stepping(global_step, boundaries, rates):
    """rates[i] applies for boundaries[i-1] <= step < boundaries[i] (ls:62-103)."""
    if any([b < 0 for b in boundaries]) or any([not isinstance(b, int) for b in boundaries]):
        raise ValueError("boundaries must be a list of positive integers")
    if any([bnext <= b for bnext, b in zip(boundaries[1:], boundaries[:-1])]):
        raise ValueError("Entries in boundaries must be strictly increasing.")
    if any([not isinstance(r, float) for r in rates]):
        raise ValueError("Learning rates must be floats")
    if len(rates) != len(boundaries) + 1:
        raise ValueError("Number of provided learning rates must exceed number of boundary points by exactly 1.")
    unreached = [i for i, b in enumerate(boundaries) if b > global_step] + [len(boundaries)]
    return rates[min(unreached)]


def from_optimizer_config(optimizer_config):
    """-> (lr_fn(step), momentum).  Only momentum_optimizer is used by the shipped configs."""
    kind = optimizer_config.WhichOneof("optimizer")
    if kind != "momentum_optimizer":
        raise ValueError("Optimizer %s not supported on the B200 path (configs use momentum_optimizer)." % kind)
    if optimizer_config.use_moving_average:
        raise ValueError("use_moving_average is not supported on the B200 path")
    cfg = optimizer_config.momentum_optimizer
    lr = cfg.learning_rate
    t = lr.WhichOneof("learning_rate")
    if t == "constant_learning_rate":
        v = float(lr.constant_learning_rate.learning_rate)
        fn = lambda step: v
    elif t == "exponential_decay_learning_rate":
        c = lr.exponential_decay_learning_rate
        def fn(step):
            e = step / float(c.decay_steps)
            if c.staircase:
                e = math.floor(e)
            return float(c.initial_learning_rate * c.decay_factor ** e)
    elif t == "manual_step_learning_rate":
        c = lr.manual_step_learning_rate
        if not c.schedule:
            raise ValueError("Empty learning rate schedule.")
        bounds = [int(x.step) for x in c.schedule]
        rates = [float(c.initial_learning_rate)] + [float(x.learning_rate) for x in c.schedule]
        fn = lambda step: manual_stepping(step, bounds, rates)
    else:
        raise ValueError("Learning_rate %s not supported." % t)
    return fn, float(cfg.momentum_optimizer_value)
