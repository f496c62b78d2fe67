"""HITRAN identifier tables the host objects need (data, not path logic): molecule short name -> HITRAN molecule
number, and molecule number -> {local isotopologue number: global isotopologue id}.

They carry exactly what the reference carries (pyradClasses.py:951-1022, the copy `Molecule.__init__` and
`getGlobalIsotope` read), row for row -- including its quirk for HBr (molecule 16), whose first two entries repeat the
ids 19 and 11 of other species -- so that `Layer.addMolecule(name_or_number, isotopeDepth)` resolves the same data
folders (`data/<global id>/`) as the reference for all 49 molecules.  tests/test_host_logic.py compares both dicts with
the reference module when it is mounted.
"""

#: HITRAN molecule numbers 1..49 in order
MOLECULE_NAMES = (
    "h2o", "co2", "o3", "n2o", "co", "ch4", "o2", "no", "so2", "no2", "nh3", "hno3", "oh", "hf", "hcl", "hbr", "hi",
    "clo", "ocs", "h2co", "hocl", "n2", "hcn", "ch3cl", "h2o2", "c2h2", "c2h6", "ph3", "cof2", "sf6", "h2s", "hcooh",
    "ho2", "o", "clono2", "no+", "hobr", "c2h4", "ch3oh", "ch3br", "ch3cn", "cf4", "c4h2", "hc3n", "h2", "cs", "so3",
    "c2n2", "cocl2")

#: global isotopologue ids per molecule number, in local-isotopologue order (index 0 = local isotopologue 1)
_GLOBAL_IDS = (
    (1, 2, 3, 4, 5, 6, 129),                            # 1 h2o
    (7, 8, 9, 10, 11, 12, 13, 14, 121, 15, 120, 122),   # 2 co2
    (16, 17, 18, 19, 20),                               # 3 o3
    (21, 22, 23, 24, 25),                               # 4 n2o
    (26, 27, 28, 29, 30, 31),                           # 5 co
    (32, 33, 34, 35),                                   # 6 ch4
    (36, 37, 38),                                       # 7 o2
    (39, 40, 41),                                       # 8 no
    (42, 43),                                           # 9 so2
    (44,),                                              # 10 no2
    (45, 46),                                           # 11 nh3
    (47, 117),                                          # 12 hno3
    (48, 49, 50),                                       # 13 oh
    (51, 110),                                          # 14 hf
    (52, 53, 107, 108),                                 # 15 hcl
    (19, 11, 111, 112),                                 # 16 hbr (sic: the reference's entries)
    (56, 113),                                          # 17 hi
    (57, 58),                                           # 18 clo
    (59, 60, 61, 62, 63),                               # 19 ocs
    (64, 65, 66),                                       # 20 h2co
    (67, 68),                                           # 21 hocl
    (69, 118),                                          # 22 n2
    (70, 71, 72),                                       # 23 hcn
    (73, 74),                                           # 24 ch3cl
    (75,),                                              # 25 h2o2
    (76, 77, 105),                                      # 26 c2h2
    (78, 106),                                          # 27 c2h6
    (79,),                                              # 28 ph3
    (80, 119),                                          # 29 cof2
    (126,),                                             # 30 sf6
    (81, 82, 83),                                       # 31 h2s
    (84,),                                              # 32 hcooh
    (85,),                                              # 33 ho2
    (86,),                                              # 34 o
    (127, 128),                                         # 35 clono2
    (87,),                                              # 36 no+
    (88, 89),                                           # 37 hobr
    (90, 91),                                           # 38 c2h4
    (92,),                                              # 39 ch3oh
    (93, 94),                                           # 40 ch3br
    (95,),                                              # 41 ch3cn
    (96,),                                              # 42 cf4
    (116,),                                             # 43 c4h2
    (109,),                                             # 44 hc3n
    (103, 115),                                         # 45 h2
    (97, 98, 99, 100),                                  # 46 cs
    (114,),                                             # 47 so3
    (123,),                                             # 48 c2n2
    (124, 125),                                         # 49 cocl2
)

MOLECULE_ID = {name: number for number, name in enumerate(MOLECULE_NAMES, start=1)}
HITRAN_GLOBAL_ISO = {number: {local: gid for local, gid in enumerate(ids, start=1)}
                     for number, ids in enumerate(_GLOBAL_IDS, start=1)}
