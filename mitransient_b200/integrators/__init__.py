from ..plugins import register_integrator
from .common import TransientADIntegrator
from .transientnlospath import CaptureType, TransientNLOSPath
from .transientpath import TransientPath

register_integrator("transient_path", lambda props: TransientPath(props))
register_integrator("transient_nlos_path", lambda props: TransientNLOSPath(props))
