"""Activation functions of jVMC/nets/activation_functions.py (:6-34) as named markers: the device kernels implement
them (csrc/cnn.cu, csrc/rbm_logpsi.cu); a marker only selects the kernel branch."""


class _Act:
    def __init__(self, name, kernel_id):
        self.__name__ = name
        self.kernel_id = kernel_id

    def __call__(self, x):
        raise NotImplementedError("%s is evaluated inside the CUDA kernels and cannot be called from Python" % self.__name__)

    def __repr__(self):
        return "<device activation %s>" % self.__name__


elu = _Act("elu", 0)
relu = _Act("relu", 1)
tanh = _Act("tanh", 2)
poly5 = _Act("poly5", 3)
poly6 = _Act("poly6", 4)
square = _Act("square", 5)
log_cosh = _Act("log_cosh", -1)      # the RBM activation (no CNN branch)

activationFunctions = {"square": square, "poly5": poly5, "poly6": poly6, "elu": elu, "relu": relu, "tanh": tanh}


def resolve(f):
    """marker, name, or a callable whose __name__ is one of the known activations -> marker"""
    if isinstance(f, _Act):
        return f
    name = f if isinstance(f, str) else getattr(f, "__name__", None)
    if name in activationFunctions:
        return activationFunctions[name]
    raise NotImplementedError("activation %r has no device kernel" % (f,))
