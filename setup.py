from setuptools import find_packages, setup

setup(
    name="blades_b200",
    version="0.1.0",
    description="Blackwell-native simulator for Byzantine-robust federated learning (API of bladesteam/blades)",
    packages=find_packages(include=["blades_b200", "blades_b200.*"]),
    package_data={"blades_b200": ["*.so", "csrc/cuda/*", "csrc/cuda/gen/*", "csrc/host/*", "csrc/*.py"]},
    python_requires=">=3.9",
    install_requires=["torch", "numpy"],
    extras_require={"data": ["torchvision"], "test": ["pytest", "hypothesis", "scikit-learn", "scipy", "pandas"]},
    zip_safe=False,
)
