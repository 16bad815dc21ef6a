from .Aligner import Aligner
