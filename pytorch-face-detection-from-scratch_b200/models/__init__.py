from .BaseModel import BaseModel
from .ModelMeta import ModelMeta
