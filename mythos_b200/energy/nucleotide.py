"""Nucleotide site transforms: interface of the reference's ``Nucleotide.from_rigid_body`` classes.

``functools.partial(Nucleotide.from_rigid_body, **geometry)`` is what the reference passes around as
``transform_fn`` (``mythos/energy/dna1/__init__.py:76-85``).  The kernels rebuild axes and sites in registers from
(center, quaternion) and only need the geometry constants, which ``Plan`` reads from the partial's keywords via
``kernel_geometry``.  Calling the transform still works: it materialises the same site arrays with torch ops on
whatever device the body lives on -- for observables and user loss functions, not for the energy path.

Formulas: ``mythos/energy/utils.py:18-36`` (axes), ``dna1/nucleotide.py:29-53``, ``dna2/nucleotide.py:30-58``,
``rna2/nucleotide.py:33-78``, ``na1/nucleotide.py:23-78``.
"""

from __future__ import annotations

import dataclasses as dc

import torch

from mythos_b200 import _lib
from mythos_b200.rigid_body import Quaternion, RigidBody


def _vec(body: RigidBody) -> torch.Tensor:
    return body.orientation.vec if isinstance(body.orientation, Quaternion) else body.orientation


def q_to_back_base(q: torch.Tensor) -> torch.Tensor:
    q0, q1, q2, q3 = q.unbind(-1)
    return torch.stack([q0**2 + q1**2 - q2**2 - q3**2, 2 * (q1 * q2 + q0 * q3), 2 * (q1 * q3 - q0 * q2)], -1)


def q_to_base_normal(q: torch.Tensor) -> torch.Tensor:
    q0, q1, q2, q3 = q.unbind(-1)
    return torch.stack([2 * (q1 * q3 + q0 * q2), 2 * (q2 * q3 - q0 * q1), q0**2 - q1**2 - q2**2 + q3**2], -1)


def q_to_cross_prod(q: torch.Tensor) -> torch.Tensor:
    q0, q1, q2, q3 = q.unbind(-1)
    return torch.stack([2 * (q1 * q2 - q0 * q3), q0**2 - q1**2 + q2**2 - q3**2, 2 * (q2 * q3 + q0 * q1)], -1)


@dc.dataclass(frozen=True)
class BaseNucleotide(RigidBody):
    stack_sites: torch.Tensor = None
    back_sites: torch.Tensor = None
    base_sites: torch.Tensor = None
    back_base_vectors: torch.Tensor = None
    base_normals: torch.Tensor = None
    cross_prods: torch.Tensor = None


def _axes(body: RigidBody):
    q = _vec(body)
    return q_to_back_base(q), q_to_base_normal(q), q_to_cross_prod(q)


def _geom(**kw) -> _lib.FlavourGeom:
    g = _lib.FlavourGeom()
    for k, v in kw.items():
        if isinstance(v, (tuple, list)):
            for i, x in enumerate(v):
                getattr(g, k)[i] = float(x)
        else:
            setattr(g, k, v)
    return g


@dc.dataclass(frozen=True)
class Dna1Nucleotide(BaseNucleotide):
    KIND = "dna1"

    @staticmethod
    def from_rigid_body(rigid_body: RigidBody, com_to_backbone, com_to_hb, com_to_stacking) -> "Dna1Nucleotide":
        a1, a3, a2 = _axes(rigid_body)
        c = rigid_body.center
        return Dna1Nucleotide(
            center=c, orientation=rigid_body.orientation, back_base_vectors=a1, base_normals=a3, cross_prods=a2,
            stack_sites=c + com_to_stacking * a1, back_sites=c + com_to_backbone * a1, base_sites=c + com_to_hb * a1,
        )

    @classmethod
    def kernel_geometry(cls, com_to_backbone, com_to_hb, com_to_stacking) -> list[_lib.FlavourGeom]:
        return [
            _geom(back=(com_to_backbone, 0.0, 0.0), stack=com_to_stacking, base=com_to_hb,
                  stack3=(com_to_stacking, 0.0), stack5=(com_to_stacking, 0.0))
        ]


@dc.dataclass(frozen=True)
class Dna2Nucleotide(BaseNucleotide):
    back_sites_dna1: torch.Tensor = None
    KIND = "dna2"

    @staticmethod
    def from_rigid_body(
        rigid_body: RigidBody, com_to_backbone_x, com_to_backbone_y, com_to_backbone_dna1, com_to_hb, com_to_stacking
    ) -> "Dna2Nucleotide":
        a1, a3, a2 = _axes(rigid_body)
        c = rigid_body.center
        return Dna2Nucleotide(
            center=c, orientation=rigid_body.orientation, back_base_vectors=a1, base_normals=a3, cross_prods=a2,
            stack_sites=c + com_to_stacking * a1, back_sites=c + com_to_backbone_x * a1 + com_to_backbone_y * a2,
            back_sites_dna1=c + com_to_backbone_dna1 * a1, base_sites=c + com_to_hb * a1,
        )

    @classmethod
    def kernel_geometry(
        cls, com_to_backbone_x, com_to_backbone_y, com_to_backbone_dna1, com_to_hb, com_to_stacking
    ) -> list[_lib.FlavourGeom]:
        return [
            _geom(back=(com_to_backbone_x, com_to_backbone_y, 0.0), back_stack=com_to_backbone_dna1, stack=com_to_stacking,
                  base=com_to_hb, stack3=(com_to_stacking, 0.0), stack5=(com_to_stacking, 0.0))
        ]


@dc.dataclass(frozen=True)
class Rna2Nucleotide(BaseNucleotide):
    bb_p3_sites: torch.Tensor = None
    bb_p5_sites: torch.Tensor = None
    stack3_sites: torch.Tensor = None
    stack5_sites: torch.Tensor = None
    KIND = "rna2"

    @staticmethod
    def from_rigid_body(
        rigid_body: RigidBody, com_to_backbone_x, com_to_backbone_y, com_to_stacking, com_to_hb,
        p3_x, p3_y, p3_z, p5_x, p5_y, p5_z, pos_stack_3_a1, pos_stack_3_a2, pos_stack_5_a1, pos_stack_5_a2,
    ) -> "Rna2Nucleotide":
        a1, a3, a2 = _axes(rigid_body)
        c = rigid_body.center
        return Rna2Nucleotide(
            center=c, orientation=rigid_body.orientation, back_base_vectors=a1, base_normals=a3, cross_prods=a2,
            back_sites=c + com_to_backbone_x * a1 + com_to_backbone_y * a3, stack_sites=c + com_to_stacking * a1,
            base_sites=c + com_to_hb * a1,
            bb_p3_sites=p3_x * a1 + p3_y * a2 + p3_z * a3, bb_p5_sites=p5_x * a1 + p5_y * a2 + p5_z * a3,
            stack3_sites=c + pos_stack_3_a1 * a1 + pos_stack_3_a2 * a2,
            stack5_sites=c + pos_stack_5_a1 * a1 + pos_stack_5_a2 * a2,
        )

    @classmethod
    def kernel_geometry(
        cls, com_to_backbone_x, com_to_backbone_y, com_to_stacking, com_to_hb,
        p3_x, p3_y, p3_z, p5_x, p5_y, p5_z, pos_stack_3_a1, pos_stack_3_a2, pos_stack_5_a1, pos_stack_5_a2,
    ) -> list[_lib.FlavourGeom]:
        return [
            _geom(back=(com_to_backbone_x, 0.0, com_to_backbone_y), stack=com_to_stacking, base=com_to_hb,
                  stack3=(pos_stack_3_a1, pos_stack_3_a2), stack5=(pos_stack_5_a1, pos_stack_5_a2),
                  p3=(p3_x, p3_y, p3_z), p5=(p5_x, p5_y, p5_z))
        ]


_DNA_KEYS = ("com_to_backbone_x", "com_to_backbone_y", "com_to_backbone_dna1", "com_to_hb", "com_to_stacking")
_RNA_KEYS = ("com_to_backbone_x", "com_to_backbone_y", "com_to_stacking", "com_to_hb", "p3_x", "p3_y", "p3_z",
             "p5_x", "p5_y", "p5_z", "pos_stack_3_a1", "pos_stack_3_a2", "pos_stack_5_a1", "pos_stack_5_a2")


@dc.dataclass(frozen=True)
class HybridNucleotide:
    dna: Dna2Nucleotide
    rna: Rna2Nucleotide
    KIND = "na1"

    @staticmethod
    def from_rigid_body(rigid_body: RigidBody, **kw) -> "HybridNucleotide":
        dna = Dna2Nucleotide.from_rigid_body(rigid_body, **{k: kw["dna_" + k] for k in _DNA_KEYS})
        rna = Rna2Nucleotide.from_rigid_body(rigid_body, **{k: kw["rna_" + k] for k in _RNA_KEYS})
        return HybridNucleotide(dna=dna, rna=rna)

    @classmethod
    def kernel_geometry(cls, **kw) -> list[_lib.FlavourGeom]:
        return [
            Dna2Nucleotide.kernel_geometry(**{k: kw["dna_" + k] for k in _DNA_KEYS})[0],
            Rna2Nucleotide.kernel_geometry(**{k: kw["rna_" + k] for k in _RNA_KEYS})[0],
        ]


for _cls in (Dna1Nucleotide, Dna2Nucleotide, Rna2Nucleotide, HybridNucleotide):
    _cls.from_rigid_body.nucleotide_cls = _cls
