from .torch_physics_trainer import Trainer, Torch_Physics_Trainer

__all__ = ["Trainer", "Torch_Physics_Trainer"]
