module Physics.B200.Decode exposing (Frame, applyState, contactPoints, frame)

{-| Elm side of the boundary, way back: what `ep_step_frame_fields` / `ep_step` returned -> the integrated bodies and the
frame's `Contacts`.  Like Encode this module lives inside the package (it rebuilds `Internal.Body.Body`) and mirrors
src/Physics.elm:581, 1210-1232 (bodies come back in the caller's list order) and 1158-1207 (`contactPoints`).

The JS half of the shim hands over plain lists read from the typed arrays; only the dynamic state travels (transform,
velocities): force / torque are zero after a step for every non-static body (src/Internal/SolverBody.elm:139-141, 221-223) and
invInertiaWorld follows from the new orientation (src/Internal/Transform3d.elm:181-205) -- both are re-derived here exactly
as `SolverBody.solved` does, so nothing differs from the reference's own result.

-}

-- `Transform3d.fromOriginAndQuaternion` / `Transform3d.quaternion`: the two accessors to add to src/Internal/Transform3d.elm
-- (shim/README.md)

import Internal.Body as Body exposing (Body)
import Internal.Shape as Shape
import Internal.Transform3d as Transform3d exposing (Transform3d)
import Internal.Vector3 as Vec3 exposing (Vec3)
import Json.Decode as D exposing (Decoder)


{-| The part of a frame that `Physics.contactPoints` reads (everything else of `Contacts` -- lambdas, friction directions,
feature keys -- stays resident on the device as the next frame's warm start): pair groups in ENUMERATION order (the reverse
of the reference's `pairGroups` list, src/Internal/BroadPhase.elm:96-109), contacts in SOLVER order (the reverse of
`pairGroup.contacts`, src/Internal/Equation.elm:135-154).
-}
type alias Frame =
    { token : Int -- frame counter kept by the shim: `simulate` passes the previous frame's back through `Contacts`
    , groupId1 : List Int
    , groupId2 : List Int
    , groupContactOff : List Int
    , groupPrePos1 : List Float -- 3 per group: origin of body1's transform3d BEFORE the step (Physics.elm:1180-1186)
    , groupPreQuat1 : List Float -- 4 per group
    , pi : List Float -- 3 per contact
    , iterations : Int
    }


frame : Decoder Frame
frame =
    D.map8 Frame
        (D.field "token" D.int)
        (D.field "group_id1" (D.list D.int))
        (D.field "group_id2" (D.list D.int))
        (D.field "group_contact_off" (D.list D.int))
        (D.field "group_pre_pos1" (D.list D.float))
        (D.field "group_pre_quat1" (D.list D.float))
        (D.field "pi" (D.list D.float))
        (D.field "iterations" D.int)


{-| Write one body's returned state (13 floats: pos 3, quat 4, vel 3, angvel 3) into the body the caller passed in;
the rest follows SolverBody.solved (src/Internal/SolverBody.elm:206-223): world shapes re-placed, invInertiaWorld from the new
orientation, force / torque cleared for non-static bodies.
-}
applyState : List Float -> Body -> Body
applyState state body =
    case state of
        [ px, py, pz, qx, qy, qz, qw, vx, vy, vz, wx, wy, wz ] ->
            if body.kindInt == 1 then
                body

            else
                let
                    transform3d =
                        Transform3d.fromOriginAndQuaternion { x = px, y = py, z = pz } qx qy qz qw
                in
                { body
                    | transform3d = transform3d
                    , velocity = { x = vx, y = vy, z = vz }
                    , angularVelocity = { x = wx, y = wy, z = wz }
                    , invInertiaWorld = Transform3d.invertedInertiaRotateIn transform3d body.invInertia
                    , worldShapesWithMaterials = List.map (\( s, m ) -> ( Shape.placeIn transform3d s, m )) body.geometry.shapesWithMaterials
                    , force = Vec3.zero
                    , torque = Vec3.zero
                }

        _ ->
            body


{-| `Physics.contactPoints` (src/Physics.elm:651-653, 1158-1207) over a decoded frame: the predicate is tried in both orders,
groups without contacts are skipped, every contact point on body1 moves from body1's pre-step frame to its post-step frame
through the composed transform `atOrigin |> relativeTo pre |> placeIn post`; the result is in pair-enumeration order (the
reference conses while walking its reversed list).  `lookup` finds a body by its ephemeral id in the post-step bodies.
-}
contactPoints : (id -> id -> Bool) -> (Int -> Maybe ( id, Body )) -> Frame -> List ( id, id, List Vec3 )
contactPoints predicate lookup f =
    let
        groups =
            List.map5 (\a b c0 c1 pre -> { id1 = a, id2 = b, c0 = c0, c1 = c1, pre = pre })
                f.groupId1
                f.groupId2
                f.groupContactOff
                (List.drop 1 f.groupContactOff)
                (chunk7 f.groupPrePos1 f.groupPreQuat1)

        points =
            chunk3 f.pi

        one g =
            if g.c1 - g.c0 <= 0 then
                Nothing

            else
                case ( lookup g.id1, lookup g.id2 ) of
                    ( Just ( ext1, body1 ), Just ( ext2, _ ) ) ->
                        if predicate ext1 ext2 || predicate ext2 ext1 then
                            let
                                transform =
                                    Transform3d.atOrigin
                                        |> Transform3d.relativeTo g.pre
                                        |> Transform3d.placeIn body1.transform3d

                                -- solver order -> `pairGroup.contacts` order
                                mine =
                                    List.reverse (List.take (g.c1 - g.c0) (List.drop g.c0 points))
                            in
                            Just ( ext1, ext2, List.map (Transform3d.pointPlaceIn transform) mine )

                        else
                            Nothing

                    _ ->
                        Nothing
    in
    List.filterMap one groups


chunk3 : List Float -> List Vec3
chunk3 xs =
    case xs of
        x :: y :: z :: rest ->
            { x = x, y = y, z = z } :: chunk3 rest

        _ ->
            []


chunk7 : List Float -> List Float -> List (Transform3d coordinates defines)
chunk7 ps qs =
    case ( ps, qs ) of
        ( px :: py :: pz :: prest, qx :: qy :: qz :: qw :: qrest ) ->
            Transform3d.fromOriginAndQuaternion { x = px, y = py, z = pz } qx qy qz qw :: chunk7 prest qrest

        _ ->
            []
