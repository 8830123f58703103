"""accelerate.utils.ProjectConfiguration (ddpm_train.py:14,254): a plain record."""
from dataclasses import dataclass
from typing import Optional


@dataclass
class ProjectConfiguration:
    project_dir: Optional[str] = None
    logging_dir: Optional[str] = None
    automatic_checkpoint_naming: bool = False
    total_limit: Optional[int] = None
    iteration: int = 0
