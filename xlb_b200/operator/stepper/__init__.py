from xlb_b200.operator.stepper.stepper import Stepper
from xlb_b200.operator.stepper.nse_stepper import IncompressibleNavierStokesStepper
