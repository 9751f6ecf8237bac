from ..plugins import register_integrator
from .common import TransientADIntegrator
from .transientnlospath import CaptureType, TransientNLOSPath
from .transientpath import TransientPath
from .transient_prbvolpath import TransientPRBVolpathIntegrator

register_integrator("transient_path", lambda props: TransientPath(props))
register_integrator("transient_nlos_path", lambda props: TransientNLOSPath(props))
register_integrator("transient_prbvolpath", lambda props: TransientPRBVolpathIntegrator(props))
