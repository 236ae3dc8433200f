"""B200-native SOM layer hot path for ViT-SOM (sm_100a, tcgen05/TMA), behind the reference's SOMLayer API."""
