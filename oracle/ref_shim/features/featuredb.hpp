// project/camera.hpp includes the feature database for a typedef it does not use on this path.
