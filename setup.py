"""Packaging for coinstac_dinunet_b200 (parity: reference setup.py:12-35).

`pip install .` installs the pure-Python package; the sm_100a kernel library is built in-tree with
`python -m coinstac_dinunet_b200.ops.build` (needs nvcc 12.8+, no GPU required) and shipped as package data."""
import pathlib

from setuptools import find_packages, setup

HERE = pathlib.Path(__file__).parent

setup(
    name='coinstac-dinunet-b200',
    version='0.1.0',
    description='Blackwell-native federated / distributed-SGD training engine with the coinstac-dinunet API.',
    long_description=(HERE / 'README.md').read_text(),
    long_description_content_type='text/markdown',
    license='MIT',
    python_requires='>=3.10',
    packages=find_packages(include=['coinstac_dinunet_b200', 'coinstac_dinunet_b200.*']),
    package_data={'coinstac_dinunet_b200.ops': ['_b200_ops.so', 'csrc/*.cu', 'csrc/*.cuh']},
    include_package_data=True,
    install_requires=['numpy', 'torch>=2.6'],
    extras_require={'vision': ['pillow', 'opencv-python-headless', 'scipy'], 'plots': ['matplotlib', 'pandas'],
                    'metrics': ['scikit-learn'], 'test': ['pytest']},
)
