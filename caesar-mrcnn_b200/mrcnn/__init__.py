"""mrcnn — B200-native drop-in for the detect path of SKA-INAF/caesar-mrcnn (mrcnn/__init__.py)."""
import logging

__title__ = "mrcnn"
__version__ = "1.0.0-b200"

logging.basicConfig(format="%(asctime)-15s %(levelname)s - %(message)s", datefmt="%Y-%m-%d %H:%M:%S")
logger = logging.getLogger(__name__)
logger.setLevel(logging.INFO)
