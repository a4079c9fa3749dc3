"""The `gpu` backend plugin set -- the third entry of pytsc's
``SIMULATOR_MODULES`` (``pytsc/__init__.py:9-14``), beside ``cityflow`` and
``sumo`` (``pytsc/backends/cityflow/__init__.py:18-26``)."""
from .config import Config, DisruptedConfig
from .metrics import MetricsParser
from .network_parser import NetworkParser
from .retriever import Retriever
from .simulator import Simulator
from .traffic_signal import TrafficSignal

GPU_MODULES = {
    "config": Config,
    "disrupted_config": DisruptedConfig,
    "metrics_parser": MetricsParser,
    "network_parser": NetworkParser,
    "retriever": Retriever,
    "simulator": Simulator,
    "traffic_signal": TrafficSignal,
}
