"""Host-side mirror of src/nn (src/nn/main.zig): activation, layer, loss and optimizer modules.  The containers
are plain glue exactly like the reference's vtables; every kernel they launch is in libwekua_b200.so."""
from . import activation, layer, loss, optimizer  # noqa: F401
from .activation import Sigmoid, Tanh  # noqa: F401
from .layer import Cache, Linear, Sequential  # noqa: F401
from .loss import mse  # noqa: F401
from .optimizer import GD, GDM, Adagrad, Adam, RMSProp  # noqa: F401
