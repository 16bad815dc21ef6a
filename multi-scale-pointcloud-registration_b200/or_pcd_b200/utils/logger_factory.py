"""Logging helper with the reference's ``LoggerFactory.get_logger`` signature
(or_pcd/utils/logger_factory.py:19-100).  One stream handler per logger name
(the reference adds a handler per call and duplicates lines), on **stderr** so that
programs printing machine-readable results on stdout (bench.py's JSON line) stay clean --
set ``OR_PCD_B200_LOG_STDOUT=1`` for the reference's stdout handler; a file handler
only when ``log_on_file`` is set AND ``OR_PCD_B200_LOG_DIR`` names a directory,
so constructing a block has no file-system side effects by default."""
import logging
import os
import sys

_FORMAT = "%(asctime)s | %(name)s | %(levelname)s : %(message)s"


class LoggerFactory:
    @staticmethod
    def get_logger(log_name: str, log_on_file: bool = False, level: int = logging.WARNING) -> logging.Logger:
        logger = logging.getLogger("or_pcd_b200." + log_name)
        if not logger.handlers:
            h = logging.StreamHandler(sys.stdout if os.environ.get("OR_PCD_B200_LOG_STDOUT") == "1" else sys.stderr)
            h.setFormatter(logging.Formatter(_FORMAT))
            logger.addHandler(h)
            logger.setLevel(level)
            logger.propagate = False
            log_dir = os.environ.get("OR_PCD_B200_LOG_DIR")
            if log_on_file and log_dir:
                os.makedirs(log_dir, exist_ok=True)
                fh = logging.FileHandler(os.path.join(log_dir, log_name))
                fh.setFormatter(logging.Formatter(_FORMAT))
                logger.addHandler(fh)
        return logger
