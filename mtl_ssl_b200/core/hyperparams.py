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
        ini = hp.initializer.WhichOneof("initializer_oneof")
        if ini == "truncated_normal_initializer":
            init = ("truncated_normal", hp.initializer.truncated_normal_initializer.stddev)
        elif ini == "variance_scaling_initializer":
            init = ("variance_scaling",)
        elif ini == "random_normal_initializer":
            init = ("normal", hp.initializer.random_normal_initializer.stddev)
        else:
            init = ("variance_scaling",)
        return Hyperparams(l2, init, hp.op, hp.activation)
