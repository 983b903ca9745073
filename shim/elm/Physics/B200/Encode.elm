module Physics.B200.Encode exposing (bodies, params, shapeTable)

{-| Elm side of the drop-in boundary (INTEGRATION.md, shim/README.md): `List ( id, Body )` after `AssignIds.assignIds`
(src/Internal/AssignIds.elm:7-27) as the flat, per-field arrays of `ep_bodies` (include/elm\_physics\_b200.h), one JSON list per
C field under the field's own name.  The JS half of the shim (patch\_simulate.js) copies each list into the typed array the
N-API addon wraps -- pure marshalling on both sides: no physics here.

This module has to live INSIDE the package (next to src/Physics.elm): it reads `Internal.Body.Body`, which the package does
not expose.  It could not be compiled in the build image (no `elm`); the byte layout it documents is exercised by
tests/test\_wire\_format.py through elm\_physics\_b200/wire.py, a line-by-line Python mirror.

Field order inside a body row follows src/Internal/Body.elm:36-61; matrices are m11,m21,m31,m12,m22,m32,m13,m23,m33
(src/Internal/Matrix3.elm:31-41); quaternions x,y,z,w (src/Internal/Transform3d.elm:379-380).

-}

import Internal.Body as Body exposing (Body)
import Internal.Material exposing (Material)
import Internal.Matrix3 exposing (Mat3)
import Internal.Shape as Shape exposing (CenterOfMassCoordinates, Shape)
import Internal.Transform3d as Transform3d
import Internal.Vector3 exposing (Vec3)
import Internal.VertexBuffer as VertexBuffer
import Json.Encode as E exposing (Value)
import Shapes.Convex as Convex exposing (Convex)


vec3 : Vec3 -> List Float
vec3 { x, y, z } =
    [ x, y, z ]


mat3 : Mat3 -> List Float
mat3 m =
    [ m.m11, m.m21, m.m31, m.m12, m.m22, m.m32, m.m13, m.m23, m.m33 ]


floats : List Float -> Value
floats =
    E.list E.float


ints : List Int -> Value
ints =
    E.list E.int


{-| One world = one `simulate` call.  `shapeIndex` maps a centre-of-mass-frame shape to its row of the shape table
(`shapeTable`; shapes are immutable after construction, src/Internal/Body.elm:64-67, so the table is built once per distinct
shape and cached by the caller).  `lin_damp_pow` / `ang_damp_pow` are NOT produced here: `(1 - damping) ^ dt` is evaluated by
the JS side with `Math.pow` (src/Internal/SolverBody.elm:149-153), so `linear_damping` / `angular_damping` travel raw.
-}
bodies : (Shape CenterOfMassCoordinates -> Int) -> List ( id, Body ) -> Value
bodies shapeIndex list =
    let
        bs =
            List.map Tuple.second list

        origin b =
            vec3 (Transform3d.originPoint b.transform3d)

        -- `Transform3d.quaternion` is one of the two accessors the package has to add to src/Internal/Transform3d.elm (the type
        -- is opaque there and `orientation` returns the rotation MATRIX, 334-345); see shim/README.md for the 8-line patch
        quat b =
            let
                q =
                    Transform3d.quaternion b.transform3d
            in
            [ q.x, q.y, q.z, q.w ]

        instCounts =
            List.map (\b -> List.length b.geometry.shapesWithMaterials) bs

        insts =
            List.concatMap (\b -> b.geometry.shapesWithMaterials) bs
    in
    E.object
        [ ( "n_worlds", E.int 1 )
        , ( "n_bodies", E.int (List.length bs) )
        , ( "world_body_off", ints [ 0, List.length bs ] )
        , ( "body_id", ints (List.map .id bs) )
        , ( "kind", ints (List.map .kindInt bs) )
        , ( "pos", floats (List.concatMap origin bs) )
        , ( "quat", floats (List.concatMap quat bs) )
        , ( "vel", floats (List.concatMap (.velocity >> vec3) bs) )
        , ( "angvel", floats (List.concatMap (.angularVelocity >> vec3) bs) )
        , ( "force", floats (List.concatMap (.force >> vec3) bs) )
        , ( "torque", floats (List.concatMap (.torque >> vec3) bs) )
        , ( "inv_inertia_world", floats (List.concatMap (.invInertiaWorld >> mat3) bs) )
        , ( "inv_mass", floats (List.map .invMass bs) )
        , ( "inv_inertia", floats (List.concatMap (.invInertia >> vec3) bs) )
        , ( "linear_damping", floats (List.map .linearDamping bs) )
        , ( "angular_damping", floats (List.map .angularDamping bs) )
        , ( "lin_lock", floats (List.concatMap (.linearLock >> vec3) bs) )
        , ( "ang_lock", floats (List.concatMap (.angularLock >> vec3) bs) )
        , ( "bounding_radius", floats (List.map (\b -> b.geometry.boundingSphereRadius) bs) )
        , ( "inst_off", ints (List.scanl (+) 0 instCounts) )
        , ( "n_insts", E.int (List.length insts) )
        , ( "inst_shape", ints (List.map (Tuple.first >> shapeIndex) insts) )
        , ( "inst_friction", floats (List.map (Tuple.second >> .friction) insts) )
        , ( "inst_bounciness", floats (List.map (Tuple.second >> .bounciness) insts) )
        ]


{-| `ep_params` of one world: units already unwrapped by `simulate` (src/Physics.elm:534-538), `collide` evaluated for every
ordered pair of the list (row-major byte matrix; omit the field when the callback is `\_ _ -> True`), `constrain` results as
rows `(a, b, kind, 24 floats)` with pivots / axes converted to centre-of-mass frames by
`Constraint.relativeToCenterOfMass` (src/Internal/Constraint.elm:16-47) -- each side with its own body's frame.
-}
params :
    { dt : Float, gravity : Vec3, iterations : Int }
    -> Maybe (List Int)
    -> List { a : Int, b : Int, kind : Int, data : List Float }
    -> Value
params cfg collide constraints =
    E.object
        ([ ( "gravity", floats (vec3 cfg.gravity) )
         , ( "dt", floats [ cfg.dt ] )
         , ( "iterations", ints [ cfg.iterations ] )
         , ( "n_constraints", E.int (List.length constraints) )
         , ( "con_body_a", ints (List.map .a constraints) )
         , ( "con_body_b", ints (List.map .b constraints) )
         , ( "con_kind", ints (List.map .kind constraints) )
         , ( "con_data", floats (List.concatMap .data constraints) )
         ]
            ++ (case collide of
                    Just bytes ->
                        [ ( "collide", ints bytes ), ( "collide_off", ints [ 0 ] ) ]

                    Nothing ->
                        []
               )
        )


{-| Rows of `ep_shapes` for a list of distinct centre-of-mass-frame shapes (kind codes of include/elm\_physics\_b200.h:
convex 0, plane 1, sphere 2, capsule 3, particle 4).  Layout of the f64 / i32 blocks: see the header.  The face-group list is
serialised AS STORED (the step walks it reversed, src/Shapes/Convex.elm:265-287); `uniqueEdges` in stored order (173).
-}
shapeTable : List (Shape CenterOfMassCoordinates) -> Value
shapeTable shapes =
    let
        rows =
            List.map shapeRow shapes

        offsets get =
            List.scanl (+) 0 (List.map (get >> List.length) rows)
    in
    E.object
        [ ( "n_shapes", E.int (List.length shapes) )
        , ( "kind", ints (List.map .kind rows) )
        , ( "f64_off", ints (offsets .f64) )
        , ( "i32_off", ints (offsets .i32) )
        , ( "f64", floats (List.concatMap .f64 rows) )
        , ( "i32", ints (List.concatMap .i32 rows) )
        ]


shapeRow : Shape CenterOfMassCoordinates -> { kind : Int, f64 : List Float, i32 : List Int }
shapeRow shape =
    case shape of
        Shape.Sphere s ->
            { kind = 2, f64 = vec3 s.position ++ [ s.radius ], i32 = [] }

        Shape.Capsule c ->
            { kind = 3, f64 = vec3 c.position ++ vec3 c.axis ++ [ c.radius, c.halfLength ], i32 = [] }

        Shape.Plane p ->
            { kind = 1, f64 = vec3 p.position ++ vec3 p.normal, i32 = [] }

        Shape.Particle p ->
            { kind = 4, f64 = vec3 p, i32 = [] }

        Shape.Convex c ->
            convexRow c


convexRow : Convex -> { kind : Int, f64 : List Float, i32 : List Int }
convexRow c =
    let
        verts =
            List.reverse (VertexBuffer.foldl (\_ v acc -> v :: acc) [] c.vertexBuffer)

        ( isBox, obb ) =
            case c.obb of
                Convex.Box ax ay az he ->
                    ( 1, vec3 ax ++ vec3 ay ++ vec3 az ++ vec3 he )

                Convex.NotBox _ _ _ _ ->
                    ( 0, List.repeat 12 0 )

        ng =
            List.length c.faces

        ne =
            List.length c.uniqueEdges

        dataOff0 =
            4 + 5 * ng + 2 * ne

        -- one pass over the face groups: 8 floats each, 5 ints each, index data appended behind all headers
        step group ( f, i, data ) =
            let
                at =
                    dataOff0 + List.length data
            in
            case group of
                Convex.OneSidedFace n idx dist _ _ _ ->
                    ( f ++ vec3 n ++ [ dist, 0, 0, 0, 0 ], i ++ [ 0, List.length idx, at, 0, 0 ], data ++ idx )

                Convex.TwoSidedFace n idx dist pn pidx pdist ->
                    ( f ++ vec3 n ++ [ dist ] ++ vec3 pn ++ [ pdist ]
                    , i ++ [ 1, List.length idx, at, List.length pidx, at + List.length idx ]
                    , data ++ idx ++ pidx
                    )

        ( groupF, groupI, faceData ) =
            List.foldl step ( [], [], [] ) c.faces

        edgeStep e ( i, data ) =
            ( i ++ [ List.length e, dataOff0 + List.length data ], data ++ e )

        ( edgeI, allData ) =
            List.foldl edgeStep ( [], faceData ) c.uniqueEdges
    in
    { kind = 0
    , f64 = vec3 c.position ++ obb ++ List.concatMap vec3 verts ++ groupF
    , i32 = [ isBox, List.length verts, ng, ne ] ++ groupI ++ edgeI ++ allData
    }
