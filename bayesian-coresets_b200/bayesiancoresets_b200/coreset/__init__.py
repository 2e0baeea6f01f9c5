from .coreset import Coreset
from .hilbert import HilbertCoreset
from .sparsevi import SparseVICoreset
from .bpsvi import BatchPSVICoreset
from .sampling import UniformSamplingCoreset
