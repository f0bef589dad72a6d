"""gs-dynamics hot paths, B200-native (sm_100a): rasterizer-with-depth + tracking iteration, GNN dynamics step."""
__version__ = "0.1.0"
