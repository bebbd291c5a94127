from .point_structure import Pointclouds
