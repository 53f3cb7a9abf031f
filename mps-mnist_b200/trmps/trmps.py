"""Import facade, as trmps/trmps.py in the reference: `from trmps import *`."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.dirname(__file__)))

from mps.mps import *
from mps.squaredDistanceMPS import *
from optimizers.optimizer import *
from optimizers.singlesiteOptimizer import SingleSiteMPSOptimizer
from optimizers.parameterObjects import *
from preprocessing.preprocessing import MPSDatasource
from preprocessing.syntheticpreprocessing import SyntheticMNISTDatasource
from preprocessing.MNISTpreprocessing import MNISTDatasource
from preprocessing.FashionMNISTpreprocessing import FashionMNISTDatasource
from trmps_native import FM_PRECOMPUTED, FM_ONE_SQRT3, FM_ONE_X, FM_COS_SIN
