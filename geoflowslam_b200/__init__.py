"""geoflowslam_b200: B200-native (sm_100a) implementation of GeoFlow-SLAM's per-frame hot path
behind the reference's own entry points.  See DESIGN.md / INTEGRATION.md."""
from ._lib import GfsError, KP_DTYPE  # noqa: F401
from .matcher import ORBmatcher  # noqa: F401
from .orb import ORBextractor  # noqa: F401
from .frontend import TrackingFrontend  # noqa: F401
from .gicp import RegistrationGICP  # noqa: F401
from .optimizer import Optimizer  # noqa: F401
from .pose import PoseOptimizer  # noqa: F401
from .pose_inertial import PoseInertialOptimizer  # noqa: F401
from .klt import KltTracker  # noqa: F401
