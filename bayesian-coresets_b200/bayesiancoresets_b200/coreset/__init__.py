from .coreset import Coreset
from .hilbert import HilbertCoreset
