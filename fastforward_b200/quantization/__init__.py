from . import affine as affine
from . import granularity as granularity
from . import tiled_tensor as tiled_tensor
