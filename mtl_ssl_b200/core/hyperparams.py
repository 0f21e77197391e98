"""What the reference's hyperparams_builder arg_scope carries for this path: the L2 weight and
the weight initializer (/root/reference/object_detection/builders/hyperparams_builder.py:24-170)."""


class Hyperparams(object):
    def __init__(self, l2_weight=0.0, init=("truncated_normal", 0.01), op="CONV", activation="RELU"):
        self.l2_weight = float(l2_weight)
        self.init = init
        self.op = op
        self.activation = activation

    @staticmethod
    def from_proto(hp):
        l2 = 0.0
        reg = hp.regularizer.WhichOneof("regularizer_oneof")
        if reg == "l2_regularizer":
            l2 = hp.regularizer.l2_regularizer.weight
        elif reg == "l1_regularizer":
            raise ValueError("l1_regularizer is not supported on the B200 path")
        # hyperparams_builder._build_initializer (hyperparams_builder.py:118-146): every parameter is passed on, an
        # unknown / missing initializer is an error
        ini = hp.initializer.WhichOneof("initializer_oneof")
        if ini == "truncated_normal_initializer":
            t = hp.initializer.truncated_normal_initializer
            init = ("truncated_normal", t.stddev, t.mean)
        elif ini == "variance_scaling_initializer":
            v = hp.initializer.variance_scaling_initializer
            mode = {0: "FAN_IN", 1: "FAN_OUT", 2: "FAN_AVG"}.get(v.mode, v.mode) if isinstance(v.mode, int) else str(v.mode)
            init = ("variance_scaling", float(v.factor), mode, bool(v.uniform))
        elif ini == "random_normal_initializer":
            r = hp.initializer.random_normal_initializer
            init = ("normal", r.stddev, r.mean)
        else:
            raise ValueError("Unknown initializer function: {}".format(ini))
        return Hyperparams(l2, init, hp.op, hp.activation)
